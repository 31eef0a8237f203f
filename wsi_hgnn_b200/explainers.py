"""Leave-one-node-out (Gem) explainer on the CUDA forward.

Reference: explainers/gem_het.py:12-41 (`HetGemExplainer`, driven by evaluator/explain_graphs.py:163-168): the causal
contribution of a node is  loss(full graph) - loss(graph without the node), one forward per node, each on a graph
rebuilt by `dgl.remove_nodes` - N + 1 sequential DGL forwards per slide.  Here the N leave-one-out graphs are built on
the device (`transforms.remove_nodes`), packed `batch` at a time into one block-diagonal graph (`hetero_graph.pack`:
its forward equals the concatenation of the per-graph forwards, trainer/train_gnn.py:59-62) and run through the same
sm_100a forward as everything else; the losses come back as one [batch] vector per launch chain.
"""
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .hetero_graph import HeteroGraph, pack
from .transforms import remove_nodes


def collapse_etypes(g: HeteroGraph, etype: str = "pos") -> HeteroGraph:
    """All relations between a pair of node types merged into ONE relation named `etype` - what gem_het.py:15-18 does by
    going through a homogeneous graph with every edge-type id zeroed.  Edge order inside the merged relation = the
    original relations in canonical order (it only affects fp summation order).  Node types are kept as they are."""
    merged: Dict = {}
    for ce in g.canonical_etypes:
        merged.setdefault((ce[0], etype, ce[2]), []).append(ce)
    edges, edata = {}, {}
    for new_ce, olds in merged.items():
        edges[new_ce] = (torch.cat([g._edges[ce][0] for ce in olds]), torch.cat([g._edges[ce][1] for ce in olds]))
        names = set.intersection(*[set(g._edata[ce].keys()) for ce in olds])
        edata[new_ce] = {n: torch.cat([g._edata[ce][n].reshape(g._edges[ce][0].shape[0], -1) for ce in olds]).squeeze(-1)
                         for n in names}
    ndata = {nt: dict(g.nodes[nt].data.items()) for nt in g.ntypes}
    return HeteroGraph({nt: g.num_nodes(nt) for nt in g.ntypes}, edges, ndata, edata)


class HetGemExplainer:
    """Same constructor and `explain_node()` contract as the reference class: node_mask[ntype][i] = loss - loss without
    node i of that type (CrossEntropyLoss against `label`)."""

    def __init__(self, graph: HeteroGraph, model, label, batch: int = 8):
        self.graph = collapse_etypes(graph, "pos")
        self.label = torch.as_tensor(label).reshape(-1).long()
        self.gnn = model
        self.batch = max(1, int(batch))

    @torch.no_grad()
    def explain_node(self, ntypes: Optional[list] = None, max_nodes: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """`ntypes` / `max_nodes` (extensions): restrict to some node types / the first max_nodes nodes of each."""
        g, dev = self.graph, self.graph.device
        label = self.label.to(dev)
        was_training = self.gnn.training
        self.gnn.eval()
        try:
            loss = F.cross_entropy(self.gnn(g), label)
            node_mask = {nt: torch.zeros(g.num_nodes(nt)) for nt in g.ntypes}
            for nt in (ntypes if ntypes is not None else g.ntypes):
                n = g.num_nodes(nt) if max_nodes is None else min(g.num_nodes(nt), max_nodes)
                for i0 in range(0, n, self.batch):
                    ids = list(range(i0, min(n, i0 + self.batch)))
                    alt = pack([remove_nodes(g, [i], nt) for i in ids])           # gem_het.py:34, `batch` graphs at once
                    pred = self.gnn(alt)                                          # [len(ids), C] = per-graph forwards
                    loss_alt = F.cross_entropy(pred, label.expand(len(ids)), reduction="none")
                    node_mask[nt][i0:i0 + len(ids)] = (loss - loss_alt).cpu()     # gem_het.py:36-38
            return node_mask
        finally:
            if was_training:
                self.gnn.train()
