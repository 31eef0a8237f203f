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
        ref = None
        for impl, tag in ((ops.IMPL_TC, "tcgen05"), (ops.IMPL_SIMT, "simt")):
            ms = timeit(lambda: ops.typed_linear(x, w, b, ptr, impl=impl, out=out, **kw), args.reps, flush)
            tf = 2.0 * N * K * n_out / (ms * 1e-3) / 1e12
            y = out.clone()
            err = None
            if ref is None:
                ref = y
            else:
                err = float((ref.double() - y.double()).norm() / y.double().norm())
            print(json.dumps({"kernel": f"typed_linear[{name}] {tag}", "N": N, "K": K, "n_out": n_out, "ms": ms,
                              "tflops": tf, "frac_bf16_peak": tf / peaks["bf16_tflops"], "rel_diff_vs_tc": err}), flush=True)
        ms = timeit(lambda: torch.matmul(x, w[0].T), args.reps, flush)
        print(json.dumps({"kernel": f"torch.matmul fp32 (cuBLAS, allow_tf32={torch.backends.cuda.matmul.allow_tf32}) [{name}]",
                          "ms": ms, "tflops": 2.0 * N * K * n_out / (ms * 1e-3) / 1e12}), flush=True)

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
