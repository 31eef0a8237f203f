"""Shared test helpers.  TEST INFRASTRUCTURE ONLY."""
import glob
import os

import torch

import golden_util
from wsi_hgnn_b200.hetero_graph import HeteroGraph, unbatch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    return sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.pt")))


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


def edge_dict_for(n_types, etypes=("neg", "pos")):
    cets = [(str(s), r, str(t)) for r in etypes for s in range(n_types) for t in range(n_types)]
    return {et: i for i, et in enumerate(cets)}


def build_oracle(model, n_types, kwargs):
    from oracle.heat import OracleHEATNet2, OracleHEATNet4
    from oracle.hgt import OracleHGT
    node_dict = {str(i): i for i in range(n_types)}
    if model == "HGT":
        return OracleHGT(node_dict, edge_dict_for(n_types), **kwargs)
    cls = {"HEATNet4": OracleHEATNet4, "HEATNet2": OracleHEATNet2}[model]
    return cls(node_dict=node_dict, **kwargs)


def build_ours(model, n_types, kwargs):
    from wsi_hgnn_b200.models import HEATNet2, HEATNet4, HGT
    node_dict = {str(i): i for i in range(n_types)}
    if model == "HGT":
        return HGT(node_dict, edge_dict_for(n_types), **kwargs)
    cls = {"HEATNet4": HEATNet4, "HEATNet2": HEATNet2}[model]
    return cls(node_dict=node_dict, **kwargs)


def golden_setup(name, builder):
    fx = load_golden(name)
    G = HeteroGraph.from_state(fx["graph"])
    m = builder(fx["model"], len(G.ntypes), fx["kwargs"])
    chk = golden_util.fill_params(m, fx["param_seed"])
    assert abs(chk - fx["param_checksum"]) <= 1e-9 * max(1.0, abs(fx["param_checksum"])), "parameter refill differs"
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert shapes == fx["param_shapes"], "state_dict keys/shapes differ from the reference's"
    m.eval()
    return fx, G, m


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def run_oracle(m, G, fx=None, independent=None):
    """Oracle logits; pack()ed graphs are the concatenation of independent per-graph forwards
    (reference trainer/train_gnn.py:59-62)."""
    if independent is None:
        independent = bool(fx.get("independent")) if fx is not None else G.independent
    with torch.no_grad():
        if independent:
            return torch.cat([m(g) for g in unbatch(G)], 0)
        return m(G)
