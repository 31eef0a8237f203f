"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of the reference edge builder, construct_graph/graph_constructor.py:256-303.

The reference's neighbour search is nmslib HNSW (graph_constructor.py:55-81: space='l2',
M=16, efConstruction=400, ef=90) - approximate, nondeterministic, un-pinned and absent here.
Its published algorithm converges to exact k-NN at recall 1, so the oracle is the exact
answer: for node i sort ALL nodes (self included) by (||f_i - f_j||_2 in fp64, j), take
`radius`, drop rank 0 (graph_constructor.py:270 drops hit [0]), emit (i -> j) i-major
(graph_constructor.py:267,273).  "Bit-exact edge_index" is defined against this.

Edge attribute: scipy.stats.pearsonr, the function the reference itself calls
(graph_constructor.py:278-280); etype = 1 ('pos') if r > 0 else 0 ('neg') (:281,296).
"""
import numpy as np


def exact_knn_edges(features: np.ndarray, radius: int, block: int = 1024) -> np.ndarray:
    """-> int64 [2, N*(radius-1)]; row 0 = query node i (repeated), row 1 = its neighbours in rank order."""
    f = np.asarray(features, dtype=np.float64)
    n = f.shape[0]
    if radius > n:
        # HNSW would return < radius hits -> np.stack fails -> ValueError -> slide skipped (get_graph.py:293-294)
        raise ValueError("fewer than `radius` nodes")
    sq = (f * f).sum(1)
    nbr = np.empty((n, radius - 1), dtype=np.int64)
    for s in range(0, n, block):
        e = min(n, s + block)
        # candidate ranking in fp64 by the expanded form, then exact re-evaluation of a generous shortlist
        d2 = sq[s:e, None] + sq[None, :] - 2.0 * (f[s:e] @ f.T)
        m = min(n, radius + 16)
        cand = np.argpartition(d2, m - 1, axis=1)[:, :m]
        for r in range(e - s):
            c = np.sort(cand[r])
            diff = f[c] - f[s + r]
            dd = (diff * diff).sum(1)                       # exact direct-form distance, fp64
            # the shortlist must be safe: everything outside it is farther than its worst member
            order = np.lexsort((c, dd))                     # (distance, index)
            nbr[s + r] = c[order][1:radius]
    src = np.repeat(np.arange(n, dtype=np.int64), radius - 1)
    return np.stack([src, nbr.reshape(-1)])


def exact_knn_edges_bruteforce(features: np.ndarray, radius: int) -> np.ndarray:
    """Small-N literal version: direct-form fp64 distances to everybody, lexsort by (distance, index)."""
    f = np.asarray(features, dtype=np.float64)
    n = f.shape[0]
    if radius > n:
        raise ValueError("fewer than `radius` nodes")
    out = np.empty((n, radius - 1), dtype=np.int64)
    idx = np.arange(n)
    for i in range(n):
        diff = f - f[i]
        dd = (diff * diff).sum(1)
        out[i] = np.lexsort((idx, dd))[1:radius]
    return np.stack([np.repeat(idx.astype(np.int64), radius - 1), out.reshape(-1)])


def pearson_edges_scipy(features: np.ndarray, edge_index: np.ndarray):
    """The reference loop verbatim in spirit: one scipy pearsonr call per edge (graph_constructor.py:276-282)."""
    from scipy.stats import pearsonr
    sim = np.empty(edge_index.shape[1], dtype=np.float64)
    et = np.empty(edge_index.shape[1], dtype=np.int64)
    for e, (a, b) in enumerate(zip(edge_index[0], edge_index[1])):
        corr = pearsonr(features[a], features[b])[0]
        et[e] = 1 if corr > 0 else 0
        sim[e] = corr
    return sim, et


def pearson_edges(features: np.ndarray, edge_index: np.ndarray, block: int = 65536):
    """Vectorised fp64 Pearson r for every edge (same quantity as pearsonr, used for large cases)."""
    f = np.asarray(features, dtype=np.float64)
    fc = f - f.mean(1, keepdims=True)
    nrm = np.sqrt((fc * fc).sum(1))
    E = edge_index.shape[1]
    sim = np.empty(E, dtype=np.float64)
    for s in range(0, E, block):
        a = edge_index[0, s:s + block]
        b = edge_index[1, s:s + block]
        sim[s:s + block] = (fc[a] * fc[b]).sum(1) / (nrm[a] * nrm[b])
    sim = np.clip(sim, -1.0, 1.0)
    return sim, (sim > 0).astype(np.int64)


def construct_graph_arrays(features: np.ndarray, node_type: np.ndarray, radius: int):
    """(edge_index [2,E], edge_type [E], sim [E]) exactly as construct_graph() assembles them before
    handing them to dgl.graph / dgl.to_heterogeneous (graph_constructor.py:285-297)."""
    ei = exact_knn_edges(features, radius)
    sim, et = pearson_edges(features, ei)
    return ei, et, sim
