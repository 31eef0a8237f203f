"""Development sweep of the kernels' debug / tuning knobs on the config-2 shapes (prints one JSON line per variant).

    python tools/sweep_dev.py [--reps 20]

GEMM: knob tc_debug bit 0 = no MMAs, bit 1 = no TMA loads, bit 2 = no epilogue body, bit 3 = no global stores, bit 4 = no
TMEM reads; every operand format (ops.set_matmul_precision).  Attention: knob attn_kernel (1 register path, 2 TMA ring,
3 TMA pipe), no_pdl.  The knobs are set through ops.dev_set (wsi_dev_set); none of them is part of the product API.
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--nodes", type=int, default=8192)
    ap.add_argument("--gemm-dbg", action="store_true", help="also time the tc_debug variants of the GEMM")
    ap.add_argument("--precisions", default="bf16x3,fp16")
    ap.add_argument("--attn-only", action="store_true")
    ap.add_argument("--k", type=int, default=5, help="out-degree of the synthetic k-NN graph of the attention sweep")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    N, T = args.nodes, 3
    if args.gemm_dbg and N > 16384:
        # the partial-pipeline debug modes (loads without MMAs, MMAs without epilogue ...) were only ever validated on the
        # config-2 shapes; at 131 072 rows one of them never finished (15 GPU-minutes lost) - keep them to small launches
        raise SystemExit("--gemm-dbg is limited to --nodes <= 16384")
    ptr = [0, N // 3, 2 * (N // 3), N]
    g = torch.Generator().manual_seed(0)

    for name, K, n_out, full in (() if args.attn_only else (("K|V|Q", 512, 1536, False), ("a_linear+skip", 512, 512, True), ("adapt_ws", 1024, 512, False))):
        x = torch.randn(N, K, generator=g).to(dev)
        w = (torch.randn(T, n_out, K, generator=g) / K ** 0.5).to(dev)
        b = torch.randn(T, n_out, generator=g).to(dev)
        kw = {}
        if full:
            kw = dict(skip=torch.ones(T, device=dev), res=torch.randn(N, n_out, device=dev),
                      row_gate=torch.ones(N, device=dev))
        tags = ((0, "full"), (4, "no epilogue body"), (8, "epilogue w/o global stores"), (16, "epilogue w/o TMEM reads"),
                (24, "epilogue: smem transposes only"), (5, "TMA only"), (6, "MMA only"), (3, "epilogue only"))
        for prec in args.precisions.split(","):
            with ops.matmul_precision(prec):
                xs, ws = ops.to_operand(x), ops.to_operand(w)
                for dbg, tag in (tags if args.gemm_dbg else tags[:1]):
                    ops.dev_set("tc_debug", dbg)
                    for want_op in (False, True):
                        ms = timeit(lambda: ops.typed_linear_op(xs, ws, b, ptr, n_out, want_op=want_op, **kw), args.reps, flush)
                        print(json.dumps({"kernel": f"typed_linear_op[{name}]", "precision": prec, "variant": tag, "y_op": want_op,
                                          "ms": ms, "tflops": 2.0 * N * K * n_out / (ms * 1e-3) / 1e12}), flush=True)
                ops.dev_set("tc_debug", 0)
    # measurement floor: an (almost) empty kernel timed the same way
    z = torch.zeros(32, device=dev)
    print(json.dumps({"kernel": "floor: z.add_(1) on 32 floats", "ms": timeit(lambda: z.add_(1.0), args.reps, flush)}), flush=True)

    D, H = 512, 4
    G = synthetic.device_slide_graph(N, 64, T, args.k, seed=1, device=dev)
    plan = G.plan()
    E = G.num_edges()
    kvq = torch.randn(N, 3 * D, device=dev)
    ew, eb = torch.ones(1, device=dev), torch.zeros(1, device=dev)
    nbytes = E * (2 * D * 4 + 8) + N * (2 * D * 4 + 4)
    work = plan.attn_work()

    def warm():
        kvq.add_(0.0)

    variants = [dict(attn_kernel=1), dict(attn_kernel=2), dict(attn_kernel=3)]
    keys = sorted({k for v in variants for k in v})
    for var in variants:
        for k in keys:
            ops.dev_set(k, var.get(k, 0))
        for op_out in (True,):
            try:
                fn = lambda: ops.hetero_attn_work(kvq[:, :D], kvq[:, D:2 * D], kvq[:, 2 * D:], work, plan.e_src, plan.e_sim,
                                                  plan.e_rel, plan.node_inv_r, ew, eb, D, H, op_out=op_out)
                ms_c = timeit(fn, args.reps, flush)
                ms_w = timeit(fn, args.reps, flush, warm)
                print(json.dumps({"kernel": "hetero_attn_work_fwd", "variant": var, "op_out": op_out, "ms_cold": ms_c,
                                  "ms_warm": ms_w, "gbs_warm": nbytes / ms_w / 1e6}), flush=True)
            except Exception as e:  # a knob combination the launch rejects
                print(json.dumps({"kernel": "hetero_attn_work_fwd", "variant": var, "error": str(e)[:200]}), flush=True)
    for k in keys:
        ops.dev_set(k, 0)


if __name__ == "__main__":
    main()
