"""Training augmentations: dgl.transforms semantics on the HeteroGraph; invariants on CPU, and on the GPU
that an augmented device graph runs through the CUDA forward and matches the oracle on the same augmented graph."""
import pytest
import torch

import golden_util
import helpers
from wsi_hgnn_b200 import synthetic
from wsi_hgnn_b200.transforms import Compose, DropEdge, DropNode, FeatMask, NodeShuffle, reference_train_transform


def _graph(seed=3):
    return synthetic.random_hetero_graph([60, 45, 0], 700, 10, seed=seed)      # one empty node type


def _edge_set(g, feats_key=True):
    """set of (relation, src feature row as tuple, dst feature row as tuple, sim) - identifies edges across relabelling"""
    out = []
    for ce in g.canonical_etypes:
        s, d = g._edges[ce]
        fs, fd = g.nodes[ce[0]].data["feat"], g.nodes[ce[2]].data["feat"]
        sim = g._edata[ce]["sim"]
        for i in range(s.numel()):
            out.append((ce, tuple(fs[s[i]].tolist()), tuple(fd[d[i]].tolist()), float(sim[i])))
    return sorted(out)


def test_drop_node_keeps_order_and_incident_edges_only():
    g = _graph()
    gen = torch.Generator().manual_seed(1)
    h = DropNode(0.4, gen)(g)
    assert h.ntypes == g.ntypes and h.canonical_etypes == g.canonical_etypes
    for nt in g.ntypes:
        a, b = g.nodes[nt].data["feat"], h.nodes[nt].data["feat"]
        assert b.shape[0] == h.num_nodes(nt) <= a.shape[0]
        # survivors keep their relative order: b is a subsequence of a
        j = 0
        for i in range(a.shape[0]):
            if j < b.shape[0] and torch.equal(a[i], b[j]):
                j += 1
        assert j == b.shape[0]
    kept_rows = {nt: {tuple(r.tolist()) for r in h.nodes[nt].data["feat"]} for nt in g.ntypes}
    want = [e for e in _edge_set(g) if e[1] in kept_rows[e[0][0]] and e[2] in kept_rows[e[0][2]]]
    assert _edge_set(h) == want                                   # exactly the edges between surviving nodes
    assert 0 < h.num_nodes() < g.num_nodes()
    h.plan()                                                      # endpoints are valid local ids


def test_drop_edge_feat_mask_node_shuffle():
    g = _graph(5)
    gen = torch.Generator().manual_seed(2)
    h = DropEdge(0.5, gen)(g)
    assert [h.num_nodes(nt) for nt in h.ntypes] == [g.num_nodes(nt) for nt in g.ntypes]
    eg, eh = _edge_set(g), _edge_set(h)
    assert 0.3 * len(eg) < len(eh) < 0.7 * len(eg) and all(e in eg for e in eh)
    big = synthetic.random_hetero_graph([30, 20], 50, 400, seed=1)
    m = FeatMask(0.5, ["feat"], gen)(big)
    for nt in big.ntypes:
        a, b = big.nodes[nt].data["feat"], m.nodes[nt].data["feat"]
        zero_cols = (b == 0).all(0)
        assert 0.35 < zero_cols.float().mean() < 0.65
        assert torch.equal(b[:, ~zero_cols], a[:, ~zero_cols])
    assert torch.equal(big.nodes["0"].data["feat"], synthetic.random_hetero_graph([30, 20], 50, 400, seed=1).nodes["0"].data["feat"])
    s = NodeShuffle(gen)(g)
    for nt in g.ntypes:
        a, b = g.nodes[nt].data["feat"], s.nodes[nt].data["feat"]
        assert sorted(map(tuple, a.tolist())) == sorted(map(tuple, b.tolist()))
    for ce in g.canonical_etypes:
        assert torch.equal(g._edges[ce][0], s._edges[ce][0]) and torch.equal(g._edges[ce][1], s._edges[ce][1])


def test_limits_and_compose():
    g = _graph(7)
    assert DropNode(0.0)(g) is g and DropEdge(0.0)(g) is g and FeatMask(0.0, ["feat"])(g) is g
    assert DropNode(1.0)(g).num_nodes() == 0 and DropEdge(1.0)(g).num_edges() == 0
    with pytest.raises(ValueError):
        DropNode(1.5)
    t = reference_train_transform(torch.Generator().manual_seed(0))
    assert isinstance(t, Compose) and [type(x).__name__ for x in t.transforms] == ["DropNode", "DropEdge", "NodeShuffle", "FeatMask"]
    h = t(g)
    assert h.num_nodes() <= g.num_nodes() and h.num_edges() <= g.num_edges()
    a = reference_train_transform(torch.Generator().manual_seed(9))(g)
    b = reference_train_transform(torch.Generator().manual_seed(9))(g)
    assert _edge_set(a) == _edge_set(b)                           # reproducible from the generator


@pytest.mark.gpu
def test_augmented_device_graph_runs_and_matches_oracle():
    dev = torch.device("cuda", 0)
    T = 3
    kw = dict(in_dim=64, hidden_dim=128, out_dim=3, n_layers=2, n_heads=4, dropuout=0.0)
    ours = helpers.build_ours("HEATNet4", T, kw)
    orc = helpers.build_oracle("HEATNet4", T, kw)
    golden_util.fill_params(ours, 3)
    orc.load_state_dict(ours.state_dict(), strict=True)
    ours, orc = ours.to(dev).eval(), orc.eval()
    G = synthetic.synth_slide_graph(1500, 64, T, 6, seed=8, noise_edges=0.1).to(dev)
    aug = reference_train_transform(torch.Generator(device=dev).manual_seed(4))(G)
    assert aug.device.type == "cuda" and 0 < aug.num_nodes() < G.num_nodes()
    with torch.no_grad():
        out = ours(aug)
        ref = orc(aug.to("cpu"))
    assert helpers.rel_err(out, ref) < 1e-3
