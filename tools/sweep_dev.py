"""Development sweep of the kernels' debug / tuning knobs on the config-2 shapes (prints one JSON line per variant).

    python tools/sweep_dev.py [--reps 20]

GEMM: WSI_TC_DEBUG bit 0 = no MMAs, bit 1 = no TMA loads, bit 2 = no epilogue body.
Attention: WSI_ATTN_RING (ring depth), WSI_ATTN_DEBUG=1 (gather 64 distinct rows), WSI_ATTN_NO_TMA (register path),
WSI_ATTN_BLOCKS (blocks per SM cap).  None of the knobs is part of the product API.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from wsi_hgnn_b200 import ops, synthetic  # noqa: E402


def timeit(fn, reps, flush, warm=None):
    """GPU time of fn: captured into a CUDA graph (so that the Python / ctypes cost of the call cannot show up as
    GPU idle time between the two events), L2 flushed (and optionally re-warmed) before every replay."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.zero_()
        if warm is not None:
            warm()
        a.record()
        g.replay()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2]


def setenv(**kw):
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--nodes", type=int, default=8192)
    ap.add_argument("--gemm-dbg", action="store_true", help="also time the WSI_TC_DEBUG variants of the GEMM")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    N, T = args.nodes, 3
    ptr = [0, N // 3, 2 * (N // 3), N]
    g = torch.Generator().manual_seed(0)

    for name, K, n_out, full in (("K|V|Q", 512, 1536, False), ("a_linear+skip", 512, 512, True), ("adapt_ws", 1024, 512, False)):
        x = torch.randn(N, K, generator=g).to(dev)
        w = (torch.randn(T, n_out, K, generator=g) / K ** 0.5).to(dev)
        b = torch.randn(T, n_out, generator=g).to(dev)
        xs, ws = ops.split_bf16(x), ops.split_bf16(w)
        kw = {}
        if full:
            kw = dict(skip=torch.ones(T, device=dev), res=torch.randn(N, n_out, device=dev),
                      row_gate=torch.ones(N, device=dev))
        tags = ((0, "full"), (4, "no epilogue body"), (8, "epilogue w/o global stores"), (16, "epilogue w/o TMEM reads"),
                (24, "epilogue: smem transposes only"), (5, "TMA only"), (6, "MMA only"), (3, "epilogue only"))
        for bn, cl in ((256, 2),):
            setenv(WSI_TC_BN=bn, WSI_TC_CL=cl)
            for dbg, tag in (tags if args.gemm_dbg else tags[:1]):
                setenv(WSI_TC_DEBUG=dbg or None)
                for want_split in (False, True):
                    ms = timeit(lambda: ops.typed_linear_split(xs, ws, b, ptr, n_out, want_split=want_split, **kw), args.reps, flush)
                    print(json.dumps({"kernel": f"typed_linear_split[{name}]", "variant": f"{tag} BN={bn} CL={cl}", "y_split": want_split,
                                      "ms": ms, "tflops": 2.0 * N * K * n_out / (ms * 1e-3) / 1e12}), flush=True)
        setenv(WSI_TC_DEBUG=None, WSI_TC_BN=None, WSI_TC_CL=None)
    # measurement floor: an (almost) empty kernel timed the same way
    z = torch.zeros(32, device=dev)
    print(json.dumps({"kernel": "floor: z.add_(1) on 32 floats", "ms": timeit(lambda: z.add_(1.0), args.reps, flush)}), flush=True)

    D, H = 512, 4
    G = synthetic.synth_slide_graph(N, 64, T, 5, seed=1).to(dev)
    plan = G.plan()
    E = G.num_edges()
    kvq = torch.randn(N, 3 * D, device=dev)
    ew, eb = torch.ones(1, device=dev), torch.zeros(1, device=dev)
    nbytes = E * (2 * D * 4 + 8) + N * (2 * D * 4 + 4)
    work = plan.attn_work()

    def warm():
        kvq.add_(0.0)

    variants = [dict(), dict(WSI_NO_PDL=1)]
    keys = sorted({k for v in variants for k in v})
    for var in variants:
        setenv(**{k: var.get(k) for k in keys})
        for split_out in (False, True):
            try:
                fn = lambda: ops.hetero_attn_work(kvq[:, :D], kvq[:, D:2 * D], kvq[:, 2 * D:], work, plan.e_src, plan.e_sim,
                                                  plan.e_rel, plan.node_inv_r, ew, eb, D, H, split_out=split_out)
                ms_c = timeit(fn, args.reps, flush)
                ms_w = timeit(fn, args.reps, flush, warm)
                print(json.dumps({"kernel": "hetero_attn_work_fwd", "variant": var, "split_out": split_out, "ms_cold": ms_c,
                                  "ms_warm": ms_w, "gbs_warm": nbytes / ms_w / 1e6}), flush=True)
            except Exception as e:  # a knob combination the launch rejects
                print(json.dumps({"kernel": "hetero_attn_work_fwd", "variant": var, "error": str(e)[:200]}), flush=True)
    setenv(**{k: None for k in keys})


if __name__ == "__main__":
    main()
