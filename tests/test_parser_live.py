"""`wsi_hgnn_b200.parser.parse_gnn_model` against the reference's OWN config files (build container only): for every YAML
under /root/reference/configs that names HGT / HEAT2 / HEAT4, the model built from the `GNN` section has the same
state_dict keys and shapes as the reference class built with the reference's constructor call (parser.py:124-171) - i.e.
`model_v{epoch}.pt` checkpoints load strictly in both directions."""
import glob
import os

import pytest
import yaml

from wsi_hgnn_b200.parser import node_and_edge_dicts, parse_gnn_model

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "configs")), reason="reference tree not present on this box")


def _configs():
    out = []
    for p in sorted(glob.glob(os.path.join(REF, "configs", "**", "*.yml"), recursive=True)):
        try:
            cfg = yaml.safe_load(open(p))
        except Exception:
            continue
        gnn = (cfg or {}).get("GNN") or {}
        if gnn.get("name") in ("HGT", "HEAT2", "HEAT4"):
            out.append((os.path.relpath(p, REF), gnn))
    return out


CONFIGS = _configs() if os.path.isdir(os.path.join(REF, "configs")) else []


def test_found_the_reference_configs():
    assert len(CONFIGS) >= 10 and {g["name"] for _, g in CONFIGS} == {"HGT", "HEAT2", "HEAT4"}


@pytest.mark.parametrize("path,gnn", CONFIGS, ids=[p for p, _ in CONFIGS])
def test_state_dict_matches_the_reference_class(path, gnn):
    import torch
    import dgl_shim
    torch.Tensor.cuda = lambda self, *a, **k: self
    mods = dgl_shim.load_reference_models(REF)
    ours = parse_gnn_model(gnn)
    if gnn["name"] == "HGT":                                                   # the reference's call, parser.py:135-143
        node_dict, edge_dict = node_and_edge_dicts(gnn["n_node_types"], gnn["edge_types"])
        ref = mods["HGT"].HGT(node_dict, edge_dict, in_dim=gnn["in_dim"], hidden_dim=gnn["hidden_dim"], out_dim=gnn["out_dim"],
                              n_layers=gnn["num_layers"], n_heads=gnn["num_heads"])
    else:                                                                      # parser.py:148-171
        name = "HEATNet2" if gnn["name"] == "HEAT2" else "HEATNet4"
        ref = getattr(mods[name], name)(in_dim=gnn["in_dim"], hidden_dim=gnn["hidden_dim"], out_dim=gnn["out_dim"],
                                        n_layers=gnn["num_layers"], n_heads=gnn["n_heads"],
                                        node_dict={str(i): i for i in range(gnn["n_node_types"])},
                                        dropuout=gnn["feat_drop"], graph_pooling_type=gnn["graph_pooling_type"])
    a = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    assert a == b
    ours.load_state_dict(ref.state_dict(), strict=True)                        # a reference checkpoint loads as is
    ref.load_state_dict(ours.state_dict(), strict=True)


def test_unknown_model_raises_like_the_reference():
    with pytest.raises(NotImplementedError):
        parse_gnn_model({"name": "GAT"})
