"""GPU parity of the edge builder (k-NN + Pearson + hetero assembly) against the CPU oracle
(oracle/knn.py: exact fp64 brute force; scipy.stats.pearsonr - the function the reference itself calls)."""
import numpy as np
import pytest
import torch

from oracle import knn as O
from wsi_hgnn_b200 import ops, synthetic
from wsi_hgnn_b200.construct_graph import GraphConstructor, Hnsw, construct_graph_arrays

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,F,radius", [(500, 64, 9), (2000, 128, 7), (33, 16, 6), (4096, 1024, 9), (2503, 128, 7), (1201, 64, 9),
                                        (1022, 256, 6)])   # incl. N % 4 != 0 on the tensor-core dot-product path
def test_knn_edges_bit_exact(n, F, radius):
    feats, _ = synthetic.synth_features(n, F, 3, seed=n)
    ref = O.exact_knn_edges(feats.numpy(), radius)
    ei, et, sim = construct_graph_arrays(feats.cuda(), radius)
    assert ei.dtype == torch.int64 and tuple(ei.shape) == (2, n * (radius - 1))
    assert np.array_equal(ei.cpu().numpy(), ref), "edge_index differs from the exact k-NN oracle"


def test_knn_small_matches_literal_bruteforce_and_duplicates():
    feats, _ = synthetic.synth_features(200, 32, 2, seed=4)
    feats[17] = feats[3]                       # exact duplicates: ties broken by index
    feats[150] = feats[3]
    ref = O.exact_knn_edges_bruteforce(feats.numpy(), 6)
    ei, _, _ = construct_graph_arrays(feats.cuda(), 6)
    assert np.array_equal(ei.cpu().numpy(), ref)


def test_knn_exact_under_adversarial_geometry():
    """The cases where a 32-candidate shortlist ranked by the fp32 expanded form ||b||^2 - 2 a.b cannot be trusted: tight
    clusters of near-duplicate patches (more than 32 points within the rounding error of the form) and features with a
    large common offset (catastrophic cancellation).  The margin check must send those rows to the exact fallback: the
    edge list still equals the literal fp64 brute force, bit for bit."""
    g = torch.Generator().manual_seed(11)
    base, _ = synthetic.synth_features(700, 64, 2, seed=9)
    # 60 near-duplicates of one patch (relative perturbation 1e-7 .. 1e-6), scattered over the index range
    dup = base[5].repeat(60, 1) * (1.0 + 1e-6 * torch.rand(60, 64, generator=g))
    pos = torch.randperm(700, generator=g)[:60]
    feats = base.clone()
    feats[pos] = dup
    ref = O.exact_knn_edges_bruteforce(feats.numpy(), 9)
    ei, _, _ = construct_graph_arrays(feats.cuda(), 9)
    assert np.array_equal(ei.cpu().numpy(), ref)
    # common offset of 300 per coordinate: ||f||^2 ~ 6e6 while neighbour distances^2 are ~ 50
    off = base + 300.0
    ref = O.exact_knn_edges_bruteforce(off.numpy(), 7)
    ei, _, _ = construct_graph_arrays(off.cuda(), 7)
    assert np.array_equal(ei.cpu().numpy(), ref)


def test_pearson_constant_row_is_nan_like_scipy():
    """scipy.stats.pearsonr of a constant row is NaN; the reference stores it and types the edge 'neg' (NaN > 0 is False,
    construct_graph/graph_constructor.py:278-281)."""
    feats, _ = synthetic.synth_features(40, 32, 2, seed=3)
    feats[7] = 1.5
    src = torch.tensor([7, 3, 7], dtype=torch.int64)
    dst = torch.tensor([2, 7, 7], dtype=torch.int64)
    sim, et = ops.edge_pearson(feats.cuda(), src.cuda(), dst.cuda())
    assert torch.isnan(sim).all() and int(et.sum()) == 0
    import warnings
    from scipy.stats import pearsonr
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert np.isnan(pearsonr(feats[7].numpy(), feats[2].numpy())[0])


def test_knn_too_few_nodes_raises():
    feats, _ = synthetic.synth_features(5, 8, 2, seed=1)
    with pytest.raises(ValueError):
        construct_graph_arrays(feats.cuda(), 9)


def test_pearson_matches_scipy():
    feats, _ = synthetic.synth_features(300, 1024, 3, seed=7)
    ei = O.exact_knn_edges(feats.numpy(), 7)
    g = np.random.default_rng(0)
    rnd = g.integers(0, 300, size=(2, 600))
    ei = np.concatenate([ei, rnd], 1)          # far pairs: correlations around 0 / negative
    sim_ref, et_ref = O.pearson_edges_scipy(feats.numpy(), ei[:, :400])
    sim64, et64 = O.pearson_edges(feats.numpy(), ei)
    sim, et = ops.edge_pearson(feats.cuda(), torch.from_numpy(ei[0]).cuda(), torch.from_numpy(ei[1]).cuda())
    sim, et = sim.cpu().numpy(), et.cpu().numpy()
    assert np.abs(sim[:400] - sim_ref).max() < 2e-6
    assert np.abs(sim - sim64).max() < 1e-6
    sure = np.abs(sim64) > 1e-6
    assert np.array_equal(et[sure], et64[sure])
    assert (et64 == 0).sum() > 50, "the test must cover the 'neg' relation"


def test_graph_constructor_end_to_end():
    n, T, radius = 600, 4, 9
    feats, ntype = synthetic.synth_features(n, 96, T, seed=11)
    het, homo, nt = GraphConstructor({"radius": radius, "n_node_type": T}, feats.numpy(), ntype.numpy()).construct_graph()
    ei, et, sim = O.construct_graph_arrays(feats.numpy(), ntype.numpy(), radius)
    assert het.ntypes == [str(t) for t in range(T)]
    assert het.num_edges() == n * (radius - 1) == homo.num_edges()
    # rebuild the homogeneous edge list from the typed graph through the _ID back-maps
    ids = {nt_: het.nodes[nt_].data["_ID"].cpu() for nt_ in het.ntypes}
    got = {}
    for ce in het.canonical_etypes:
        s, d = het.edges(etype=ce)
        for a, b, r in zip(ids[ce[0]][s.cpu()].tolist(), ids[ce[2]][d.cpu()].tolist(),
                           het.edata["sim"][ce].cpu().tolist() if isinstance(het.edata["sim"], dict) else het.edata["sim"].cpu().tolist()):
            got[(a, b)] = (ce[1], r)
    assert len(got) == ei.shape[1]
    for e in range(ei.shape[1]):
        name, r = got[(int(ei[0, e]), int(ei[1, e]))]
        assert abs(r - sim[e]) < 1e-6
        if abs(sim[e]) > 1e-6:
            assert name == ("pos" if et[e] else "neg")


def test_hnsw_query_surface():
    feats, _ = synthetic.synth_features(100, 16, 2, seed=2)
    m = Hnsw(space="l2").fit(feats.numpy())
    all_nbr = m.query_all(5).cpu().numpy()
    one = m.query(feats[10].numpy(), 5)
    assert list(one) == list(all_nbr[10])


@pytest.mark.parametrize("seed,n,T,hub", [(0, 3000, 3, 0), (1, 500, 6, 300), (2, 20000, 2, 0)])
def test_native_plan_matches_host_plan(seed, n, T, hub):
    """wsi_plan_build_csr / wsi_plan_attn_work_* == the host (torch ops) statement of the same layout: bit-exact."""
    import torch
    from wsi_hgnn_b200 import synthetic
    if hub:
        G = synthetic.random_hetero_graph([n // T] * T, 4 * n, 8, seed=seed, hub=hub)
    else:
        G = synthetic.synth_slide_graph(n, 16, T, 5, seed=seed, noise_edges=0.3)
    host = G.plan()                                   # CPU graph -> host path
    Gd = G.to("cuda")
    dev = Gd.plan()                                   # CUDA graph -> plan-builder kernels
    assert dev._stats is not None
    for name in ("rowptr", "e_src", "e_rel", "e_sim"):
        assert torch.equal(getattr(dev, name).cpu(), getattr(host, name)), name
    for chunk in (4, 16):
        wd, wh = dev.attn_work(chunk), host.attn_work(chunk)
        assert (wd["n_items"], wd["n_split"], wd["n_part"]) == (wh["n_items"], wh["n_split"], wh["n_part"])
        di, hi = wd["items"][:wd["n_items"]].cpu(), wh["items"]
        npart = wh["n_part"]
        assert torch.equal(di[:npart], hi[:npart])                       # chunk items: edge order, exact
        assert torch.equal(di[npart:, 2] - di[npart:, 1], hi[npart:, 2] - hi[npart:, 1])   # same size sequence
        key = lambda x: x[torch.argsort(x[:, 0])]                        # order inside a size class is arbitrary
        assert torch.equal(key(di[npart:]), key(hi[npart:]))
        if wh["n_split"]:
            assert torch.equal(wd["split_row"].cpu(), wh["split_row"])
            assert torch.equal(wd["split_ptr"].cpu(), wh["split_ptr"])
            assert torch.equal(wd["part_rel"].cpu(), wh["part_rel"])
            assert torch.equal(wd["part_split"].cpu(), wh["part_split"])
    assert dev.max_in_degree == host.max_in_degree


def test_native_plan_rejects_out_of_range_edges():
    import torch
    from wsi_hgnn_b200.hetero_graph import HeteroGraph
    G = HeteroGraph({"0": 4, "1": 3}, {("0", "pos", "1"): (torch.tensor([0, 1, 3]), torch.tensor([0, 7, 2]))},
                    {"0": {"feat": torch.zeros(4, 8)}, "1": {"feat": torch.zeros(3, 8)}})
    with pytest.raises(IndexError):
        p = G.to("cuda").plan()
        p.check()
