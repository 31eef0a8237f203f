"""Leave-one-node-out explainer: host-side graph surgery on CPU, and on the GPU the batched explainer
against the literal one-forward-per-node loop of the reference run on the oracle."""
import pytest
import torch
import torch.nn.functional as F

import golden_util
import helpers
from wsi_hgnn_b200 import synthetic
from wsi_hgnn_b200.explainers import HetGemExplainer, collapse_etypes
from wsi_hgnn_b200.transforms import remove_nodes


def test_collapse_etypes_and_remove_nodes():
    g = synthetic.random_hetero_graph([12, 9], 80, 6, seed=2)
    c = collapse_etypes(g)
    assert all(ce[1] == "pos" for ce in c.canonical_etypes) and c.num_edges() == g.num_edges()
    assert len(c.canonical_etypes) == len({(ce[0], ce[2]) for ce in g.canonical_etypes})
    for ce in c.canonical_etypes:
        olds = [o for o in g.canonical_etypes if (o[0], o[2]) == (ce[0], ce[2])]
        assert torch.equal(c._edges[ce][0], torch.cat([g._edges[o][0] for o in olds]))
        assert torch.equal(c._edata[ce]["sim"], torch.cat([g._edata[o]["sim"] for o in olds]))
    r = remove_nodes(c, [3], "0")
    assert r.num_nodes("0") == 11 and r.num_nodes("1") == 9 and r.canonical_etypes == c.canonical_etypes
    assert torch.equal(r.nodes["0"].data["feat"], torch.cat([c.nodes["0"].data["feat"][:3], c.nodes["0"].data["feat"][4:]]))
    for ce in c.canonical_etypes:
        s, d = c._edges[ce]
        m = torch.ones_like(s, dtype=torch.bool)
        if ce[0] == "0":
            m &= s != 3
        if ce[2] == "0":
            m &= d != 3
        fix = lambda x, t: x - (x > 3).long() if t == "0" else x
        assert torch.equal(r._edges[ce][0], fix(s[m], ce[0])) and torch.equal(r._edges[ce][1], fix(d[m], ce[2]))


@pytest.mark.gpu
def test_batched_explainer_matches_literal_loop_on_the_oracle():
    dev = torch.device("cuda", 0)
    T = 2
    kw = dict(in_dim=16, hidden_dim=128, out_dim=2, n_layers=2, n_heads=4, dropuout=0.0)
    ours = helpers.build_ours("HEATNet4", T, kw)
    orc = helpers.build_oracle("HEATNet4", T, kw)
    golden_util.fill_params(ours, 13)
    orc.load_state_dict(ours.state_dict(), strict=True)
    ours, orc = ours.to(dev).eval(), orc.eval()
    G = synthetic.random_hetero_graph([14, 11], 120, 16, seed=6)
    label = torch.tensor([1])
    mask = HetGemExplainer(G.to(dev), ours, label, batch=5).explain_node()
    # the reference's loop, literally (explainers/gem_het.py:25-41), on the CPU oracle
    c = collapse_etypes(G)
    with torch.no_grad():
        loss = F.cross_entropy(orc(c), label)
        for nt in c.ntypes:
            ref = torch.tensor([float(loss - F.cross_entropy(orc(remove_nodes(c, [i], nt)), label))
                                for i in range(c.num_nodes(nt))])
            assert mask[nt].shape == ref.shape
            assert torch.allclose(mask[nt], ref, atol=2e-5, rtol=1e-3), (nt, (mask[nt] - ref).abs().max())
