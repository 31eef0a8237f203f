"""Synthetic WSI slide graphs of TCGA shape (SURVEY.md §8d "Configs -> concrete synthetic inputs").

features = mixture of 8T Gaussian clusters in R^F (centres N(0,1), points = centre + 0.5 N(0,1));
node type = cluster id mod T; edges = exact k-NN in feature space (k = radius-1, query -> neighbour),
sim = Pearson r, etype = r > 0 ('pos') else 'neg' - the quantities the reference builder produces
(construct_graph/graph_constructor.py:256-303).  There is no network, so no real TCGA data.

The host (torch, CPU) k-NN here is only the data generator for tests and benches; the product
builder is wsi_hgnn_b200.construct_graph (CUDA).
"""
from typing import Optional, Sequence

import torch

from .hetero_graph import HeteroGraph, to_heterogeneous

TYPE_SKEW6 = [.45, .30, .15, .05, .03, .02]


def synth_features(n: int, f: int, n_types: int, seed: int, skew: bool = False, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    n_clusters = 8 * n_types
    centres = torch.randn(n_clusters, f, generator=g)
    if skew and n_types == 6:
        probs = torch.tensor(TYPE_SKEW6).repeat(8) / 8.0          # cluster c has type c % 6
        cid = torch.multinomial(probs, n, replacement=True, generator=g)
    else:
        cid = torch.randint(0, n_clusters, (n,), generator=g)
    feats = centres[cid] + 0.5 * torch.randn(n, f, generator=g)
    ntype = cid % n_types
    return feats.to(device), ntype.to(device)


def host_knn(feats: torch.Tensor, k: int, block: int = 2048) -> torch.Tensor:
    """Exact k-NN on the host: fp64 expanded-form ranking with (distance, index) order, self dropped."""
    f = feats.double()
    n = f.shape[0]
    sq = (f * f).sum(1)
    out = torch.empty(n, k, dtype=torch.int64)
    for s in range(0, n, block):
        e = min(n, s + block)
        d2 = sq[s:e, None] + sq[None, :] - 2.0 * (f[s:e] @ f.T)
        d2[torch.arange(e - s), torch.arange(s, e)] = -1.0        # self is rank 0
        idx = torch.topk(d2, k + 1, dim=1, largest=False, sorted=True).indices
        out[s:e] = idx[:, 1:]
    return out


def pearson(feats: torch.Tensor, src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    f = feats.double()
    fc = f - f.mean(1, keepdim=True)
    nrm = fc.norm(dim=1)
    r = (fc[src] * fc[dst]).sum(1) / (nrm[src] * nrm[dst])
    return r.clamp(-1, 1).float()


def synth_slide_graph(n: int, f: int, n_types: int, k: int, seed: int, skew: bool = False,
                      etypes: Sequence[str] = ("neg", "pos"), noise_edges: float = 0.0) -> HeteroGraph:
    """One slide graph on the CPU.  `noise_edges` > 0 rewires that fraction of edges to random
    destinations (gives the 'neg' relations some population for tests)."""
    feats, ntype = synth_features(n, f, n_types, seed, skew)
    nbr = host_knn(feats, k)
    src = torch.arange(n).repeat_interleave(k)
    dst = nbr.reshape(-1)
    if noise_edges > 0:
        g = torch.Generator().manual_seed(seed + 7919)
        m = torch.rand(src.numel(), generator=g) < noise_edges
        dst = torch.where(m, torch.randint(0, n, (src.numel(),), generator=g), dst)
    sim = pearson(feats, src, dst)
    et = (sim > 0).long()
    return to_heterogeneous(src, dst, ntype, et, [str(t) for t in range(n_types)], list(etypes),
                            ndata={"feat": feats}, edata={"sim": sim})


def random_hetero_graph(num_nodes: Sequence[int], n_edges: int, f: int, seed: int,
                        etypes: Sequence[str] = ("neg", "pos"), hub: Optional[int] = None) -> HeteroGraph:
    """Unstructured random typed graph (random edges, random sim in [-1,1]) for parity tests.
    `hub`: if given, that many extra edges all point at node 0 (in-degree >> 32 case)."""
    g = torch.Generator().manual_seed(seed)
    T = len(num_nodes)
    n = int(sum(num_nodes))
    ntype = torch.cat([torch.full((c,), t, dtype=torch.int64) for t, c in enumerate(num_nodes)])
    ntype = ntype[torch.randperm(n, generator=g)]
    src = torch.randint(0, n, (n_edges,), generator=g)
    dst = torch.randint(0, n, (n_edges,), generator=g)
    if hub:
        src = torch.cat([src, torch.randint(0, n, (hub,), generator=g)])
        dst = torch.cat([dst, torch.zeros(hub, dtype=torch.int64)])
    sim = torch.rand(src.numel(), generator=g) * 2 - 1
    et = torch.randint(0, len(etypes), (src.numel(),), generator=g)
    feats = torch.randn(n, f, generator=g)
    return to_heterogeneous(src, dst, ntype, et, [str(t) for t in range(T)], list(etypes),
                            ndata={"feat": feats}, edata={"sim": sim})


def device_slide_graph(n: int, f: int, n_types: int, k: int, seed: int, device, skew: bool = False) -> HeteroGraph:
    """The same recipe built ON THE GPU (device RNG; the product's own k-NN + Pearson kernels as the edge builder): the
    generator of the large benchmark batches (256 slides of 2k-20k nodes), where the exact fp64 host k-NN of
    synth_slide_graph would take minutes per slide.  Different random stream than the host generator."""
    from .construct_graph.graph_constructor import construct_graph_arrays
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    n_clusters = 8 * n_types
    centres = torch.randn(n_clusters, f, generator=g, device=dev)
    if skew and n_types == 6:
        probs = torch.tensor(TYPE_SKEW6, device=dev).repeat(8) / 8.0
        cid = torch.multinomial(probs, n, replacement=True, generator=g)
    else:
        cid = torch.randint(0, n_clusters, (n,), generator=g, device=dev)
    feats = centres[cid] + 0.5 * torch.randn(n, f, generator=g, device=dev)
    ntype = cid % n_types
    ei, et, sim = construct_graph_arrays(feats, k + 1)
    return to_heterogeneous(ei[0], ei[1], ntype, et.to(torch.int64), [str(t) for t in range(n_types)], ["neg", "pos"],
                            ndata={"feat": feats}, edata={"sim": sim})
