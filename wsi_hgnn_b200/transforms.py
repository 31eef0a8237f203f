"""Training augmentations of the reference on the DEVICE-resident graph.

The reference applies `dgl.transforms.Compose([DropNode(0.5), DropEdge(0.5), NodeShuffle(), FeatMask(0.5, ['feat'])])`
to every training sample on the CPU (data.py:16-23, applied at :116-117, :222-223, :281-282), which changes the graph
every step and puts the CSR build on the critical path.  These classes keep the `dgl.transforms` names and semantics
[DGL-mem] but run as a handful of torch ops on whatever device the HeteroGraph lives on; the plan (CSR + work list) of
the augmented graph is rebuilt by the device-side planner on first use.  Random draws come from torch (an optional
`generator`), not from DGL's RNG: parity with the reference is distributional, not per-sample.

  DropNode(p)    every node is removed with probability p, together with its incident edges; the survivors keep their
                 relative order and are relabelled 0..n'-1 per node type
  DropEdge(p)    every edge is removed with probability p (per canonical edge type)
  NodeShuffle()  the node feature rows of every node type are permuted among its nodes (structure unchanged)
  FeatMask(p, node_feat_names)  every feature COLUMN of the named node features is zeroed with probability p
"""
from typing import Dict, List, Optional, Sequence

import torch

from .hetero_graph import HeteroGraph


def _rand(n: int, device, generator: Optional[torch.Generator]):
    if generator is not None and generator.device != torch.device(device):
        return torch.rand(n, generator=generator, device=generator.device).to(device)
    return torch.rand(n, generator=generator, device=device)


class BaseTransform:
    def __call__(self, g: HeteroGraph) -> HeteroGraph:
        raise NotImplementedError


class Compose(BaseTransform):
    """dgl.transforms.Compose: apply the transforms in order."""

    def __init__(self, transforms: Sequence[BaseTransform]):
        self.transforms = list(transforms)

    def __call__(self, g: HeteroGraph) -> HeteroGraph:
        for t in self.transforms:
            g = t(g)
        return g


def _rebuild(g: HeteroGraph, num_nodes: Dict[str, int], edges, ndata, edata) -> HeteroGraph:
    if g.batch_size != 1:
        raise ValueError("augmentations apply to single slides (before batching / packing), as in the reference's Dataset")
    return HeteroGraph(num_nodes, edges, ndata, edata)


def keep_nodes(g: HeteroGraph, keep: Dict[str, torch.Tensor]) -> HeteroGraph:
    """Node-induced subgraph: `keep[nt]` bool [n_nt] (node types not listed keep all their nodes).  Survivors keep their
    relative order and are relabelled 0..n'-1 per node type; edges with a removed endpoint disappear; every node type
    and relation of `g` stays in the result, possibly empty (dgl.remove_nodes semantics [DGL-mem])."""
    dev = g.device
    new_id: Dict[str, torch.Tensor] = {}
    counts: List[torch.Tensor] = []
    masks: Dict[str, torch.Tensor] = {}
    for nt in g.ntypes:
        k = keep.get(nt)
        if k is None:
            k = torch.ones(g.num_nodes(nt), dtype=torch.bool, device=dev)
        masks[nt] = k
        new_id[nt] = torch.cumsum(k.to(torch.int64), 0) - 1                # valid where k
        counts.append(k.sum())
    num_nodes, ndata = {}, {}
    for nt, c in zip(g.ntypes, torch.stack(counts).tolist() if counts else []):        # one host read for all types
        num_nodes[nt] = int(c)
        ndata[nt] = {name: v[masks[nt]] for name, v in g.nodes[nt].data.items()}
    edges, edata = {}, {}
    for ce in g.canonical_etypes:
        s, d = g._edges[ce]
        m = masks[ce[0]][s] & masks[ce[2]][d] if s.numel() else torch.zeros(0, dtype=torch.bool, device=dev)
        edges[ce] = (new_id[ce[0]][s[m]], new_id[ce[2]][d[m]])
        edata[ce] = {name: v[m] for name, v in g._edata[ce].items()}
    return _rebuild(g, num_nodes, edges, ndata, edata)


def remove_nodes(g: HeteroGraph, nids, ntype: str) -> HeteroGraph:
    """dgl.remove_nodes(g, nids, ntype=...) (explainers/gem_het.py:34)."""
    k = torch.ones(g.num_nodes(ntype), dtype=torch.bool, device=g.device)
    k[torch.as_tensor(nids, device=g.device).reshape(-1).long()] = False
    return keep_nodes(g, {ntype: k})


class DropNode(BaseTransform):
    def __init__(self, p: float = 0.5, generator: Optional[torch.Generator] = None):
        if not 0.0 <= p <= 1.0:
            raise ValueError("p must be in [0, 1]")
        self.p, self.generator = p, generator

    def __call__(self, g: HeteroGraph) -> HeteroGraph:
        if self.p == 0:
            return g
        return keep_nodes(g, {nt: _rand(g.num_nodes(nt), g.device, self.generator) >= self.p for nt in g.ntypes})


class DropEdge(BaseTransform):
    def __init__(self, p: float = 0.5, generator: Optional[torch.Generator] = None):
        if not 0.0 <= p <= 1.0:
            raise ValueError("p must be in [0, 1]")
        self.p, self.generator = p, generator

    def __call__(self, g: HeteroGraph) -> HeteroGraph:
        if self.p == 0:
            return g
        dev = g.device
        edges, edata = {}, {}
        for ce in g.canonical_etypes:
            s, d = g._edges[ce]
            m = _rand(int(s.numel()), dev, self.generator) >= self.p
            edges[ce] = (s[m], d[m])
            edata[ce] = {name: v[m] for name, v in g._edata[ce].items()}
        ndata = {nt: dict(g.nodes[nt].data.items()) for nt in g.ntypes}
        return _rebuild(g, {nt: g.num_nodes(nt) for nt in g.ntypes}, edges, ndata, edata)


class NodeShuffle(BaseTransform):
    def __init__(self, generator: Optional[torch.Generator] = None):
        self.generator = generator

    def __call__(self, g: HeteroGraph) -> HeteroGraph:
        dev = g.device
        ndata = {}
        for nt in g.ntypes:
            n = g.num_nodes(nt)
            perm = torch.argsort(_rand(n, dev, self.generator))
            ndata[nt] = {name: v[perm] for name, v in g.nodes[nt].data.items()}
        edges = {ce: g._edges[ce] for ce in g.canonical_etypes}
        edata = {ce: dict(g._edata[ce].items()) for ce in g.canonical_etypes}
        return _rebuild(g, {nt: g.num_nodes(nt) for nt in g.ntypes}, edges, ndata, edata)


class FeatMask(BaseTransform):
    def __init__(self, p: float = 0.5, node_feat_names: Optional[Sequence[str]] = None,
                 generator: Optional[torch.Generator] = None):
        if not 0.0 <= p <= 1.0:
            raise ValueError("p must be in [0, 1]")
        self.p, self.names, self.generator = p, (list(node_feat_names) if node_feat_names else []), generator

    def __call__(self, g: HeteroGraph) -> HeteroGraph:
        if self.p == 0 or not self.names:
            return g
        dev = g.device
        ndata = {}
        for nt in g.ntypes:
            fr = dict(g.nodes[nt].data.items())
            for name in self.names:
                if name in fr and fr[name].dim() == 2 and fr[name].shape[0] > 0:
                    colkeep = (_rand(int(fr[name].shape[1]), dev, self.generator) >= self.p).to(fr[name].dtype)
                    fr[name] = fr[name] * colkeep.unsqueeze(0)             # (a new tensor: the stored slide is not modified)
            ndata[nt] = fr
        edges = {ce: g._edges[ce] for ce in g.canonical_etypes}
        edata = {ce: dict(g._edata[ce].items()) for ce in g.canonical_etypes}
        return _rebuild(g, {nt: g.num_nodes(nt) for nt in g.ntypes}, edges, ndata, edata)


def reference_train_transform(generator: Optional[torch.Generator] = None) -> Compose:
    """The transform the reference hard-codes for its training split (data.py:16-23)."""
    return Compose([DropNode(0.5, generator), DropEdge(0.5, generator), NodeShuffle(generator),
                    FeatMask(0.5, ["feat"], generator)])
