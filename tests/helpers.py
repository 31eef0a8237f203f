"""Shared test helpers.  TEST INFRASTRUCTURE ONLY."""
import contextlib
import glob
import os

import torch
import torch.nn.functional as F

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


@contextlib.contextmanager
def oracle_rounding(model, dtype=torch.bfloat16, store_kv: bool = True, min_k: int = 64, min_out: int = 64):
    """Run the CPU oracle with the product's 16-bit rounding at the product's storage points (set_matmul_precision("bf16") /
    ("fp16")): every nn.Linear the tensor-core path takes (in_features >= 64 and % 8 == 0, out_features >= 64 - the
    [B, *] readout heads stay fp32) sees its input and its weight rounded to `dtype`, accumulates in fp32, adds the fp32
    bias; with `store_kv` the outputs of the k_linears / v_linears are additionally stored in `dtype` (the bf16-storage
    configuration of BASELINE config 3).  TEST INFRASTRUCTURE ONLY."""
    kv = set()
    if store_kv:
        for layer in getattr(model, "gcs", []):
            for tag in ("k_linears", "v_linears"):
                for lin in getattr(layer, tag, []):
                    kv.add(id(lin.weight))
    orig = F.linear

    def lin(x, w, b=None):
        if x.dim() != 2 or w.shape[1] < min_k or w.shape[1] % 8 != 0 or w.shape[0] < min_out:
            return orig(x, w, b)
        y = torch.mm(x.float().to(dtype).float(), w.float().to(dtype).float().t())
        if b is not None:
            y = y + b.float()
        if id(w) in kv:
            y = y.to(dtype).float()
        return y.to(x.dtype)
    F.linear = torch.nn.functional.linear = lin
    try:
        yield
    finally:
        F.linear = torch.nn.functional.linear = orig
