"""Randomised live cross-check of the oracle against the reference's own model files (build container only: needs
/root/reference; skipped on the GPU box).  Beyond the committed fixtures: every run draws fresh small graphs - varying
node-type counts, empty types, hubs, isolated nodes, missing relations, all poolings - executes the UNMODIFIED
models/HEATNet4.py, HEATNet2.py and HGT.py on the DGL stand-in (tests/dgl_shim.py) in fp64 and requires the oracle to
agree to 1e-9.  This is what pins the oracle to the reference's code rather than to a reading of it."""
import os

import pytest
import torch

import golden_util
import helpers
from wsi_hgnn_b200 import synthetic
from wsi_hgnn_b200.hetero_graph import batch

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference tree not present on this box")


def _ref_models():
    import dgl_shim
    torch.Tensor.cuda = lambda self, *a, **k: self          # models/HEATNet4.py:240 hard-codes .cuda()
    return dgl_shim, dgl_shim.load_reference_models("/root/reference")


def _fp64(m):
    m = m.double()
    for mod in m.modules():
        if hasattr(mod, "e_linear"):
            mod.e_linear.float()                            # the reference casts sim to fp32 (models/HEATNet4.py:103)
    return m


CASES = []
for seed in range(4):
    g = torch.Generator().manual_seed(100 + seed)
    T = int(torch.randint(1, 5, (1,), generator=g))
    sizes = torch.randint(0, 14, (T,), generator=g).tolist()
    sizes[int(torch.randint(0, T, (1,), generator=g))] = 9                     # at least one populated type
    n_edges = int(torch.randint(5, 90, (1,), generator=g))
    hub = int(torch.randint(0, 40, (1,), generator=g)) if seed % 2 else None
    for model, pooling in (("HEATNet4", "mean"), ("HEATNet2", "max"), ("HEATNet4", "sum"), ("HGT", "mean"), ("HGT", "sum")):
        CASES.append((seed, T, tuple(sizes), n_edges, hub, model, pooling))


@pytest.mark.parametrize("seed,T,sizes,n_edges,hub,model,pooling", CASES)
def test_oracle_equals_reference_on_random_graphs(seed, T, sizes, n_edges, hub, model, pooling):
    dgl_shim, mods = _ref_models()
    G = synthetic.random_hetero_graph(list(sizes), n_edges, 12, seed=200 + seed, hub=hub)
    if seed == 3:                                                               # DGL-batch semantics too
        G2 = synthetic.random_hetero_graph(list(sizes), n_edges + 7, 12, seed=300 + seed, hub=hub)
        if G2.canonical_etypes == G.canonical_etypes:
            G = batch([G, G2])
    node_dict = {str(i): i for i in range(T)}
    if model == "HGT":
        kw = dict(in_dim=12, hidden_dim=24, out_dim=3, n_layers=2, n_heads=4, use_norm=bool(seed % 2), graph_pooling_type=pooling)
        ref = mods["HGT"].HGT(node_dict, helpers.edge_dict_for(T), **kw)
    else:
        kw = dict(in_dim=12, hidden_dim=24, out_dim=3, n_layers=2, n_heads=4, dropuout=0.2, graph_pooling_type=pooling)
        ref = getattr(mods[model], model)(node_dict=node_dict, **kw)
    golden_util.fill_params(ref, 77 + seed)
    orc = helpers.build_oracle(model, T, kw)
    orc.load_state_dict(ref.state_dict(), strict=True)                           # same keys and shapes as the reference
    ref, orc = _fp64(ref.eval()), _fp64(orc.eval())
    for nt in G.ntypes:
        G.nodes[nt].data["feat"] = G.nodes[nt].data["feat"].double()
    sg = dgl_shim.shim_graph_from(G)
    with torch.no_grad():
        want = ref(sg)
        got = orc(G)
    assert got.shape == want.shape
    assert helpers.rel_err(got, want) < 1e-9, (model, pooling, sizes)
