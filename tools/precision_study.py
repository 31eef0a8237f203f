"""Which operand precision can the tensor-core typed GEMM (K1) use and stay inside the 1e-3 parity bar?

CPU experiment (no GPU): the oracle's nn.Linear calls are replaced by emulations of the candidate tensor-core
schemes (operands rounded exactly as the hardware would see them, products accumulated in fp64 and rounded to fp32 -
at least as accurate as the fp32 accumulator in TMEM) and the whole forward is compared with the fp64 oracle on the
16 reference-generated goldens and on BASELINE config 2 at full size.  The error reported is the worse of the logits
and node-embedding relative L2 errors.  Output: profiles/r2_precision_study.json (a table in DESIGN.md section 3).

    python tools/precision_study.py [--no-config2]
"""
import argparse
import contextlib
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def tf32_trunc(x):
    return (x.float().contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def tf32_rn(x):
    i = x.float().contiguous().view(torch.int32)
    i = i + 0xFFF + ((i >> 13) & 1)
    return (i & ~0x1FFF).view(torch.float32)


def split(x, dt):
    hi = x.float().to(dt).float()
    lo = (x.float() - hi).to(dt).float()
    return hi, lo


def mm(a, w):
    return (a.double() @ w.double().t()).float()


_ORIG_LINEAR = F.linear


def scheme_fp32(x, w):
    return _ORIG_LINEAR(x.float(), w.float())


def scheme_tf32_trunc(x, w):
    return mm(tf32_trunc(x), tf32_trunc(w))


def scheme_tf32_rn(x, w):
    return mm(tf32_rn(x), tf32_rn(w))


def make_split3(dt):
    def f(x, w):
        xh, xl = split(x, dt)
        wh, wl = split(w, dt)
        return (xh.double() @ wh.double().t() + xh.double() @ wl.double().t() + xl.double() @ wh.double().t()).float()
    return f


def make_split2_a_hi(dt):
    """A rounded to one term, W kept in two: A_hi*W_hi + A_hi*W_lo (2 passes)."""
    def f(x, w):
        xh, _ = split(x, dt)
        wh, wl = split(w, dt)
        return (xh.double() @ (wh.double() + wl.double()).t()).float()
    return f


def make_split2_w_hi(dt):
    def f(x, w):
        xh, xl = split(x, dt)
        wh, _ = split(w, dt)
        return ((xh.double() + xl.double()) @ wh.double().t()).float()
    return f


def scheme_single(dt):
    def f(x, w):
        return mm(x.float().to(dt).float(), w.float().to(dt).float())
    return f


SCHEMES = {
    "fp32 (oracle arithmetic)": (scheme_fp32, 0),
    "1xTF32 truncating (hardware kind::tf32 on raw fp32)": (scheme_tf32_trunc, 2),
    "1xTF32 round-to-nearest operands": (scheme_tf32_rn, 2),
    "1xbf16": (scheme_single(torch.bfloat16), 1),
    "1xfp16": (scheme_single(torch.float16), 1),
    "2xfp16 A_hi*(W_hi+W_lo)": (make_split2_a_hi(torch.float16), 2),
    "2xfp16 (A_hi+A_lo)*W_hi": (make_split2_w_hi(torch.float16), 2),
    "2xbf16 A_hi*(W_hi+W_lo)": (make_split2_a_hi(torch.bfloat16), 2),
    "1xfp16, K|V stored fp16 (Q fp32)": ((scheme_single(torch.float16), (torch.float16, "kv")), 1),
    "1xfp16, K|V|Q stored fp16": ((scheme_single(torch.float16), (torch.float16, "kvq")), 1),
    "1xbf16, K|V|Q stored bf16 (config-3 storage)": ((scheme_single(torch.bfloat16), (torch.bfloat16, "kvq")), 1),
    "3xbf16 hi*hi+hi*lo+lo*hi (shipped)": (make_split3(torch.bfloat16), 3),
    "3xfp16 hi*hi+hi*lo+lo*hi": (make_split3(torch.float16), 3),
}


KVQ_WEIGHTS = {}          # id(weight) -> 'k' | 'v' | 'q' of the model under test


@contextlib.contextmanager
def patched_linear(fn):
    orig = F.linear
    store = None
    if isinstance(fn, tuple):
        fn, store = fn

    def lin(x, w, b=None):
        if w.shape[1] < 32 or x.dim() != 2 or x.shape[0] < 8:      # e_linear (1x1), the [B,*] heads: fp32 SIMT in the product
            return orig(x, w, b)
        y = fn(x, w).to(x.dtype)
        y = y if b is None else y + b
        if store is not None and KVQ_WEIGHTS.get(id(w), "") in store[1]:
            y = y.to(store[0]).to(x.dtype)                            # K|V|Q kept in 16-bit storage for the edge phase
        return y
    F.linear = lin
    torch.nn.functional.linear = lin
    try:
        yield
    finally:
        F.linear = orig
        torch.nn.functional.linear = orig


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def run_case(name, G, m32, m64, independent):
    from wsi_hgnn_b200.hetero_graph import unbatch
    import copy

    def fwd(m, dbl):
        outs, embs = [], []
        gs = unbatch(G) if independent else [G]
        for g in gs:
            if dbl:
                g = g.double_features() if hasattr(g, "double_features") else g
            o, h = m(g, return_embeddings=True)
            outs.append(o)
            embs.append(torch.cat([h[nt] for nt in g.ntypes if h[nt].shape[0] > 0], 0))
        return torch.cat(outs, 0), torch.cat(embs, 0)

    res = {}
    KVQ_WEIGHTS.clear()
    for layer in m32.gcs:
        for tag in "kvq":
            for lin_m in getattr(layer, tag + "_linears"):
                KVQ_WEIGHTS[id(lin_m.weight)] = tag
    with torch.no_grad():
        ref_o, ref_h = fwd(m64, True)
        for sname, (fn, _) in SCHEMES.items():
            with patched_linear(fn):
                o, h = fwd(m32, False)
            res[sname] = max(rel(o, ref_o), rel(h, ref_h))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-config2", action="store_true")
    args = ap.parse_args()
    import copy
    import helpers
    torch.set_num_threads(os.cpu_count())
    table = {}
    cases = [c for c in helpers.golden_cases()]
    for c in cases:
        fx, G, m = helpers.golden_setup(c, helpers.build_oracle)
        m64 = copy.deepcopy(m).double()
        G64 = G
        try:
            table[c] = run_case(c, G, m, _Dbl(m64), bool(fx.get("independent")))
        except TypeError:
            continue                                   # model without return_embeddings (HGT oracle): logits only
        print(c, {k: f"{v:.2e}" for k, v in table[c].items()}, flush=True)
    if not args.no_config2:
        sys.argv = [sys.argv[0]]
        import bench
        G = bench.make_graph(1)
        _, orc = bench.build_models(True, False)
        m64 = copy.deepcopy(orc).double()
        table["config2"] = run_case("config2", G, orc, _Dbl(m64), False)
        print("config2", {k: f"{v:.2e}" for k, v in table["config2"].items()}, flush=True)
    worst = {s: max(t[s] for t in table.values()) for s in SCHEMES}
    out = {"tolerance": 1e-3, "margin_required": 3.0,
           "schemes": [{"scheme": s, "bf16_rate_passes": SCHEMES[s][1], "worst_rel_err": worst[s],
                        "config2_rel_err": table.get("config2", {}).get(s),
                        "margin": (1e-3 / worst[s]) if worst[s] > 0 else None} for s in SCHEMES],
           "cases": table}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "r2_precision_study.json"), "w") as f:
        json.dump(out, f, indent=1)
    for r in out["schemes"]:
        print(f"{r['scheme']:58s} passes={r['bf16_rate_passes']} worst={r['worst_rel_err']:.2e} margin={r['margin']:.1f}x")


class _Dbl:
    """fp64 oracle: features and sim are promoted on the fly."""

    def __init__(self, m):
        self.m = m

    def __call__(self, g, **kw):
        g2 = _promote(g)
        orig = F.linear
        F.linear = torch.nn.functional.linear = lambda x, w, b=None: orig(x.to(w.dtype), w, b)
        try:
            return self.m(g2, **kw)
        finally:
            F.linear = torch.nn.functional.linear = orig


def _promote(g):
    import copy
    g2 = copy.copy(g)
    from wsi_hgnn_b200.hetero_graph import HeteroGraph
    st = g.state()
    def conv(o):
        if isinstance(o, torch.Tensor) and o.is_floating_point():
            return o.double()
        if isinstance(o, dict):
            return {k: conv(v) for k, v in o.items()}
        if isinstance(o, (list, tuple)):
            return type(o)(conv(v) for v in o)
        return o
    return HeteroGraph.from_state(conv(st))


if __name__ == "__main__":
    main()
