"""Per-kernel timings of the hot-path kernels on the config-2 shapes (CUDA events, L2 flushed between reps).

    python tools/bench_kernels.py [--reps 20]

Prints one JSON line per kernel: ms (median), achieved TFLOP/s or GB/s against MEASURED_PEAKS.json.
Development tool; bench.py is the contract benchmark.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from wsi_hgnn_b200 import ops, synthetic  # noqa: E402


def timeit(fn, reps, flush):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--nodes", type=int, default=8192)
    ap.add_argument("--scale-nodes", type=int, default=163840,
                    help="rows of the batch-scale GEMM rows (a packed training micro-batch, T=6 skewed types); 0 = skip")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    N, T = args.nodes, 3
    ptr = [0, N // 3, 2 * (N // 3), N]
    g = torch.Generator().manual_seed(0)

    def gemm_case(name, K, n_out, **epi):
        x = torch.randn(N, K, generator=g).to(dev)
        w = (torch.randn(T, n_out, K, generator=g) / K ** 0.5).to(dev)
        b = torch.randn(T, n_out, generator=g).to(dev)
        out = torch.empty(N, n_out, device=dev)
        kw = {}
        if epi.get("skip"):
            kw = dict(skip=torch.ones(T, device=dev), res=torch.randn(N, n_out, device=dev),
                      row_gate=torch.ones(N, device=dev))
        ref64 = torch.cat([x[ptr[t]:ptr[t + 1]].double() @ w[t].double().T + b[t].double() for t in range(T)]) if not kw else None
        flops = 2.0 * N * K * n_out
        for prec in ("fp16", "bf16x3", "bf16"):
            with ops.matmul_precision(prec):
                xs, ws = ops.to_operand(x), ops.to_operand(w)
                # the GEMM alone on operands already in operand form (what the forward chain launches) ...
                ms = timeit(lambda: ops.typed_linear_op(xs, ws, b, ptr, n_out, **kw), args.reps, flush)
                y, _ = ops.typed_linear_op(xs, ws, b, ptr, n_out, **kw)
                err = float((ref64 - y.double()).norm() / ref64.norm()) if ref64 is not None else None
                issued = {"fp16": 1, "bf16": 1, "bf16x3": 3}[prec]
                tf = flops / (ms * 1e-3) / 1e12
                print(json.dumps({"kernel": f"typed_linear_op[{name}] tcgen05 {prec}", "N": N, "K": K, "n_out": n_out, "ms": ms,
                                  "tflops": tf, "frac_bf16_peak": tf / peaks["bf16_tflops"],
                                  "frac_issued": issued * tf / peaks["bf16_tflops"], "rel_err_vs_fp64": err}), flush=True)
                # ... and through the fp32 API (conversion pre-pass of x and w inside the timed region)
                ms = timeit(lambda: ops.typed_linear(x, w, b, ptr, impl=ops.IMPL_TC, out=out, **kw), args.reps, flush)
                print(json.dumps({"kernel": f"typed_linear[{name}] tcgen05 {prec} incl. fp32 -> operand pre-pass", "ms": ms,
                                  "tflops": flops / (ms * 1e-3) / 1e12}), flush=True)
        ms = timeit(lambda: ops.typed_linear(x, w, b, ptr, impl=ops.IMPL_SIMT, out=out, **kw), args.reps, flush)
        print(json.dumps({"kernel": f"typed_linear[{name}] simt fp32", "ms": ms, "tflops": flops / (ms * 1e-3) / 1e12}), flush=True)
        # library bars on the same box (one dense [N, K] x [K, n_out] product, no epilogue, no type grouping)
        wt = w[0].T.contiguous()
        for tag, setup in (("cuBLAS fp32 (allow_tf32=False)", lambda: setattr(torch.backends.cuda.matmul, "allow_tf32", False)),
                           ("cuBLASLt TF32 (allow_tf32=True)", lambda: setattr(torch.backends.cuda.matmul, "allow_tf32", True))):
            setup()
            ms = timeit(lambda: torch.matmul(x, wt), args.reps, flush)
            err = float(((x.double() @ wt.double()) - torch.matmul(x, wt).double()).norm() / (x.double() @ wt.double()).norm())
            print(json.dumps({"kernel": f"torch.matmul {tag} [{name}]", "ms": ms, "tflops": flops / (ms * 1e-3) / 1e12,
                              "rel_err_vs_fp64": err}), flush=True)
        torch.backends.cuda.matmul.allow_tf32 = False
        # 3xTF32-class: torch's "high" float32 matmul precision (cuBLASLt picks its own split scheme when it has one)
        try:
            torch.set_float32_matmul_precision("high")
            ms = timeit(lambda: torch.matmul(x, wt), args.reps, flush)
            print(json.dumps({"kernel": f"torch.matmul float32_matmul_precision=high [{name}]", "ms": ms,
                              "tflops": flops / (ms * 1e-3) / 1e12}), flush=True)
        finally:
            torch.set_float32_matmul_precision("highest")
        for dt, tag in ((torch.float16, "fp16"), (torch.bfloat16, "bf16")):
            xa, wa = x.to(dt), wt.to(dt)
            ms = timeit(lambda: torch.matmul(xa, wa), args.reps, flush)
            print(json.dumps({"kernel": f"torch.matmul cuBLASLt {tag} in / {tag} out [{name}]", "ms": ms,
                              "tflops": flops / (ms * 1e-3) / 1e12}), flush=True)

    def scale_case(Ns):
        """The same kernels at the size of a packed training micro-batch (config 5: ~160k nodes, 6 skewed types): the
        config-2 launches above are 3 tiles per CTA pair - prologue and epilogue of a 25 us kernel - these are ~50."""
        Ts = 6
        cnt = [int(Ns * f) for f in synthetic.TYPE_SKEW6]
        cnt[0] += Ns - sum(cnt)
        tp = [0]
        for c in cnt:
            tp.append(tp[-1] + c)
        for name, K, n_out in (("K|V|Q", 512, 1536), ("a_linear", 512, 512), ("adapt_ws", 1024, 512)):
            x = torch.randn(Ns, K, device=dev)
            w = torch.randn(Ts, n_out, K, device=dev) / K ** 0.5
            b = torch.randn(Ts, n_out, device=dev)
            dy = torch.randn(Ns, n_out, device=dev)
            flops = 2.0 * Ns * K * n_out
            for prec in ("fp16", "bf16x3"):
                with ops.matmul_precision(prec):
                    xs, ws = ops.to_operand(x), ops.to_operand(w)
                    ms = timeit(lambda: ops.typed_linear_op(xs, ws, b, tp, n_out), args.reps, flush)
                issued = 3 if prec == "bf16x3" else 1
                tf = flops / (ms * 1e-3) / 1e12
                print(json.dumps({"kernel": f"typed_linear_op[{name}] tcgen05 {prec} (batch scale)", "N": Ns, "K": K, "n_out": n_out,
                                  "ms": ms, "tflops": tf, "frac_bf16_peak": tf / peaks["bf16_tflops"],
                                  "frac_issued": issued * tf / peaks["bf16_tflops"]}), flush=True)
            xs, ds = ops.to_operand(x, ops.OPF_BF16X3), ops.to_operand(dy, ops.OPF_BF16X3)
            ms = timeit(lambda: ops.typed_wgrad(ds, xs, tp), args.reps, flush)
            tf = flops / (ms * 1e-3) / 1e12
            print(json.dumps({"kernel": f"typed_wgrad[{name}] tcgen05 bf16x3 MN-major + reduce (batch scale)", "N": Ns, "M": n_out,
                              "Nn": K, "ms": ms, "tflops": tf, "frac_bf16_peak": tf / peaks["bf16_tflops"],
                              "frac_issued": 3 * tf / peaks["bf16_tflops"]}), flush=True)
            xh, dh = xs[:Ns], ds[:Ns]

            def cublas3():
                for t in range(Ts):
                    a_, z_ = tp[t], tp[t + 1]
                    torch.mm(dh[a_:z_].t(), xh[a_:z_], out_dtype=torch.float32)
            try:
                ms1 = timeit(cublas3, args.reps, flush)
                print(json.dumps({"kernel": f"cuBLASLt bf16 -> fp32 dY^T X per type, ONE of the three terms [{name}] (batch scale)",
                                  "ms": ms1, "tflops_one_term": flops / (ms1 * 1e-3) / 1e12}), flush=True)
            except Exception as e:       # noqa: BLE001 - library bar only
                print(json.dumps({"kernel": f"cuBLASLt wgrad bar [{name}]", "error": str(e)[:100]}), flush=True)

    if args.scale_nodes > 0:
        scale_case(args.scale_nodes)

    gemm_case("K|V|Q", 512, 1536)
    gemm_case("a_linear+skip", 512, 512, skip=True)
    gemm_case("adapt_ws", 1024, 512)

    # edge attention on the config-2 graph
    D, H = 512, 4
    G = synthetic.synth_slide_graph(N, 64, T, 5, seed=1).to(dev)
    plan = G.plan()
    E = G.num_edges()
    kvq = torch.randn(N, 3 * D, device=dev)
    ew, eb = torch.ones(1, device=dev), torch.zeros(1, device=dev)
    use_perm = ops.head_perm(D, H) is not None
    fn = lambda: ops.hetero_attn(kvq[:, :D], kvq[:, D:2 * D], kvq[:, 2 * D:], plan.rowptr, plan.e_src, plan.e_sim,
                                 plan.e_rel, plan.node_inv_r, ew, eb, D, H, use_perm)
    ms = timeit(fn, args.reps, flush)
    nbytes = E * (2 * D * 4 + 8) + N * (2 * D * 4 + 4)
    print(json.dumps({"kernel": "hetero_attn_fwd whole rows (cold L2)", "E": E, "ms": ms, "gbs": nbytes / ms / 1e6,
                      "frac_hbm_peak": nbytes / ms / 1e6 / peaks["hbm_gbs"]}), flush=True)
    for chunk in (8, 16, 32):
        work = plan.attn_work(chunk)
        fn = lambda: ops.hetero_attn_work(kvq[:, :D], kvq[:, D:2 * D], kvq[:, 2 * D:], work, plan.e_src, plan.e_sim,
                                          plan.e_rel, plan.node_inv_r, ew, eb, D, H)
        ms = timeit(fn, args.reps, flush)
        print(json.dumps({"kernel": f"hetero_attn_work_fwd chunk={chunk} (cold L2)", "n_items": work["n_items"],
                          "n_part": work["n_part"], "ms": ms, "gbs": nbytes / ms / 1e6,
                          "frac_hbm_peak": nbytes / ms / 1e6 / peaks["hbm_gbs"]}), flush=True)

        def warm():
            kvq.add_(0.0)
        # K/V/Q left in L2 by the producing GEMM (the in-model situation)
        for _ in range(3):
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.reps)]
        for a_, b_ in ev:
            flush.zero_(); warm(); a_.record(); fn(); b_.record()
        torch.cuda.synchronize()
        msw = sorted(a_.elapsed_time(b_) for a_, b_ in ev)[len(ev) // 2]
        print(json.dumps({"kernel": f"hetero_attn_work_fwd chunk={chunk} (K|V|Q warm in L2)", "ms": msw,
                          "gbs": nbytes / msw / 1e6, "frac_hbm_peak": nbytes / msw / 1e6 / peaks["hbm_gbs"]}), flush=True)
    ms = timeit(fn, args.reps, flush)
    print(json.dumps({"kernel": "hetero_attn_work_fwd last (cold L2)", "E": E, "ms": ms, "gbs": nbytes / ms / 1e6,
                      "frac_hbm_peak": nbytes / ms / 1e6 / peaks["hbm_gbs"]}), flush=True)


if __name__ == "__main__":
    main()
