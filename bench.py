"""bench.py - HEATNet4 forward throughput (edges/s) on the BASELINE.json config-2 workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one HEATNet4 forward (input projection + 3 HEAT layers + typed readout + heads) over one synthetic
TCGA-BRCA-shape slide graph (8192 patch nodes, 3 node types, 40 960 edges, F=1024, D=512, H=4, fp32).
  value     : edges/s with the graph resident in HBM (CSR pre-built), CUDA-graph replay, CUDA-event timed,
              L2 flushed between steps
  e2e       : the same metric through the public API from pinned HOST buffers (slide_io.stream_forward): per slide
              one H2D copy of features + edge arrays, CSR build, forward, logits D2H - all inside the timed region
  roofline  : the edge-attention kernel (gather K/V by source, per-relation softmax, scatter to dst), HBM bound
  cpu_baseline / --impl reference : the reference-structured CPU restatement (oracle/) on the host cores; the
              reference itself cannot run here (DGL absent)
N > 1 (torchrun): every rank runs its own slides (independent units, no collective in the forward): weak scaling.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CFG = dict(workload="config2: TCGA-BRCA-shape synthetic slide graph, 3-layer HEATNet4 forward, fp32",
           nodes=8192, node_types=3, k=5, edges=40960, in_dim=1024, hidden=512, heads=4, layers=3, out_dim=2,
           graphs_per_step=1)
METRIC = "HEATNet4 fwd edges/sec"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"


def ncu_traffic(which: str):
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(which)
    return None


def make_graph(seed: int):
    from wsi_hgnn_b200 import synthetic
    return synthetic.synth_slide_graph(CFG["nodes"], CFG["in_dim"], CFG["node_types"], CFG["k"], seed=seed)


def model_kwargs():
    T = CFG["node_types"]
    return dict(in_dim=CFG["in_dim"], hidden_dim=CFG["hidden"], out_dim=CFG["out_dim"], n_layers=CFG["layers"],
                n_heads=CFG["heads"], node_dict={str(i): i for i in range(T)}, dropuout=0.2)


def build_models(want_oracle: bool, want_ours: bool):
    import golden_util
    ours = orc = None
    if want_oracle:
        from oracle.heat import OracleHEATNet4
        torch.manual_seed(611)
        orc = OracleHEATNet4(**model_kwargs())
        golden_util.fill_params(orc, 611)
        orc.eval()
    if want_ours:
        from wsi_hgnn_b200.models import HEATNet4
        torch.manual_seed(611)
        ours = HEATNet4(**model_kwargs())
        golden_util.fill_params(ours, 611)
        ours.eval()
    return ours, orc


class ClockSampler:
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md clocks line).  NVML in-process (one sample
    per millisecond: the timed region is only tens of ms long); nvidia-smi as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None
        self.th = threading.Thread(target=self.run, daemon=True)

    def sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        bits = [getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8), getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20), getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)]
        return [str(sm), str(mx)] + ["Active" if r & b else "Not Active" for b in bits]

    def run(self):
        while not self.stop.is_set():
            try:
                if self.nvml is not None:
                    self.rows.append(self.sample_nvml())
                    self.stop.wait(0.001)
                    continue
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                self.nvml = None
            self.stop.wait(0.1)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def cpu_forward_time(orc, G, reps: int, warm: int):
    times = []
    with torch.no_grad():
        for i in range(warm + reps):
            t0 = time.perf_counter()
            orc(G)
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
    times.sort()
    return times[len(times) // 2]


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path = the oracle port (oracle/), all host threads."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    _, orc = build_models(True, False)
    G = make_graph(1)
    E = G.num_edges()
    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 2))):
            orc(G)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            orc(G)
        dt = time.perf_counter() - t0
    v = E * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "edges/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(CFG, note="reference-structured PyTorch CPU restatement (oracle/), not DGL"),
            "cpu_baseline": {"value": v, "unit": "edges/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} full config-2 forwards"},
            "e2e": {"value": v, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from wsi_hgnn_b200 import _lib, ops
    from wsi_hgnn_b200.graphed import GraphedForward
    from wsi_hgnn_b200.hetero_graph import HeteroGraph

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    # torchrun pins OMP_NUM_THREADS=1; the synthetic slides are generated on the host (exact fp64 k-NN), so give every
    # rank its share of the cores for that set-up phase
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // max(world, 1)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL printf()s its version banner to stdout when NCCL_DEBUG is VERSION / WARN; stdout must carry ONE JSON line:
        # send NCCL's log to stderr and keep fd 1 pointed at stderr while the communicator is created
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import ctypes
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            try:
                ctypes.CDLL(None).fflush(None)
            except Exception:
                pass
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    lib = _lib.load()
    hbm_peak, tc_peak, peak_src = peaks()

    ours, orc = build_models(rank == 0 and not args.no_cpu_baseline, True)
    ours = ours.to(dev)
    G_host = make_graph(1 + rank)                    # every rank has its own slide (weak scaling)
    E, N, D, L = G_host.num_edges(), G_host.num_nodes(), CFG["hidden"], CFG["layers"]
    G = G_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    # ---------------------------------------------------------------- resident-input forward (value)
    gf = GraphedForward(ours, G, warmup=args.warmup)
    for _ in range(args.warmup):
        gf()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(local) as clk:
        for a, b in ev:
            flush.zero_()                            # L2 flush between timed iterations (not timed)
            a.record()
            gf()
            b.record()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_dev = sum(a.elapsed_time(b) for a, b in ev) * 1e-3
    t = torch.tensor([t_dev], device=dev, dtype=torch.float64)
    e_tot = torch.tensor([float(E)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(e_tot, op=dist.ReduceOp.SUM)
    t_max, edges_all = float(t), float(e_tot)
    value = edges_all * args.steps / t_max

    # eager (Python-issued launches, no CUDA graph) for comparison
    with torch.no_grad():
        for _ in range(3):
            ours(G)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            ours(G)
        b.record()
        torch.cuda.synchronize()
    eager_ms = a.elapsed_time(b) / 10

    # ---------------------------------------------------------------- end to end from host buffers (e2e)
    # public API: slide_io.FlatSlide (one pinned blob per slide) + slide_io.stream_forward (H2D of slide i+1 on the
    # copy engine overlaps the forward of slide i; logits come back through pinned memory).  Every slide is copied
    # host -> device, planned (CSR + work list) and run; nothing is reused between slides.
    from wsi_hgnn_b200.slide_io import FlatSlide, stream_forward
    n_e2e = max(8, min(args.steps, 40))
    distinct = [FlatSlide.from_graph(G_host if i == 0 else make_graph(101 + 7 * rank + i), pin=True) for i in range(4)]
    slides = [distinct[i % len(distinct)] for i in range(n_e2e)]
    h2d = sum(s_.header["nbytes"] for s_ in slides) // n_e2e
    e2e_edges = torch.tensor([float(sum(s_.num_edges() for s_ in slides))], device=dev, dtype=torch.float64)
    for _ in range(2):
        list(stream_forward(ours, slides[:4], dev))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    outs, stamps = [], []
    for o_ in stream_forward(ours, slides, dev):
        outs.append(o_)
        stamps.append(time.perf_counter())
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if os.environ.get("WSI_BENCH_DEBUG"):
        gaps = [stamps[0] - t0] + [b_ - a_ for a_, b_ in zip(stamps, stamps[1:])]
        print("e2e per-slide gaps (ms):", " ".join(f"{1e3 * g_:.2f}" for g_ in gaps), file=sys.stderr, flush=True)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_edges, op=dist.ReduceOp.SUM)
    e2e_value = float(e2e_edges) / float(t_e2e)
    d2h = outs[0].numel() * 4
    # the same slides one at a time with a host sync per slide (what the reference's evaluation loop does)
    for s_ in slides[:3]:
        with torch.no_grad():
            ours(s_.to_graph(dev, non_blocking=True)).cpu()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s_ in slides[:8]:
        with torch.no_grad():
            ours(s_.to_graph(dev, non_blocking=True)).cpu()
    e2e_sync_ms = (time.perf_counter() - t0) / 8 * 1e3

    # ---------------------------------------------------------------- roofline of the edge-attention kernel
    layer = ours.gcs[0]
    plan = G.plan()
    with torch.no_grad():
        x = torch.randn(N, D, device=dev)
        order = [ours.node_dict[nt] for nt in plan.ntypes]
        w_kvq, b_kvq, wa, ba, skip, use_perm = layer._packed(order)
        w_kvq_s, wa_s = layer._packed_split(order)
        xs = ops.to_operand(x)
        kvq, _ = ops.typed_linear_op(xs, w_kvq_s, b_kvq, plan.type_ptr, 3 * D)
        work = plan.attn_work()
        attn_args = (kvq[:, :D], kvq[:, D:2 * D], kvq[:, 2 * D:], work, plan.e_src, plan.e_sim, plan.e_rel,
                     plan.node_inv_r, layer.e_linear.weight, layer.e_linear.bias, D, CFG["heads"])
        agg_out = torch.empty(N, D, device=dev)

        def graphed(fn):
            """capture one call of fn: the replay costs no Python / ctypes time between the timing events"""
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):
                fn()
            return g_

        g_attn = graphed(lambda: ops.hetero_attn_work(*attn_args, out=agg_out))
        reps = 20
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in kev:
            flush.zero_()
            kvq.add_(0.0)                             # K/V/Q back in L2 as the producing GEMM leaves them
            a.record()
            g_attn.replay()
            b.record()
        torch.cuda.synchronize()
        attn_ms = sorted(a.elapsed_time(b) for a, b in kev)[reps // 2]
        # dense: fused K|V|Q typed GEMM on pre-split operands (as inside the forward)
        g_gemm = graphed(lambda: ops.typed_linear_op(xs, w_kvq_s, b_kvq, plan.type_ptr, 3 * D))
        gev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in gev:
            flush.zero_()
            a.record()
            g_gemm.replay()
            b.record()
        torch.cuda.synchronize()
        gemm_ms = sorted(a.elapsed_time(b) for a, b in gev)[reps // 2]
    attn_bytes = E * (2 * D * 4 + 8) + N * (2 * D * 4 + 4)          # SURVEY.md §8(d) per-layer edge-phase bytes
    achieved = attn_bytes / (attn_ms * 1e-3) / 1e9
    gemm_flops = 2.0 * N * D * 3 * D
    gemm_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    line = {"metric": METRIC, "value": value, "unit": "edges/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_max / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(CFG, l2="flushed between timed iterations (256 MB write)", launch="CUDA-graph replay",
                           eager_ms_per_step=eager_ms, parallelism=f"dp{world} (independent slides per rank)"),
            "e2e": {"value": e2e_value, "unit": "edges/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(t_e2e) / n_e2e * 1e3, "steps": n_e2e,
                    "sync_ms_per_step": e2e_sync_ms,
                    "note": "slide_io.stream_forward over pinned FlatSlide blobs: per slide ONE H2D copy (features + edges + "
                            "sim), CSR + work-list build and forward issued by one C call (wsi_slide_forward), logits D2H; the "
                            "copy of slide i+1 overlaps the forward of slide i.  sync_ms_per_step = the same slides one at a "
                            "time through model(G) with a host sync per slide"},
            "gpu_launches": gf.kernels_per_replay * args.steps,
            "roofline": {"kernel": "edge attention of one layer (wsi_hetero_attn_work_fwd: TMA bulk-copy gather of K|V by "
                                   "source, per-relation softmax, merge of hub-row chunks, write to dst)", "bound": "hbm",
                         "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": ncu_traffic("attn"), "peak_source": peak_src, "algorithmic_bytes": attn_bytes,
                         "kernel_ms": attn_ms, "note": "algorithmic bytes = SURVEY 8(d) edge-phase bytes (every gathered "
                         "K/V row counted); K|V (33.5 MB) is L2 resident, so DRAM traffic is far below it"},
            "roofline_dense": {"kernel": "typed_linear_tc_kernel K|V|Q [8192x512]x[512x1536], 3-term bf16 split "
                                         "(fp32-accurate) on tcgen05", "bound": "tensor",
                               "achieved": gemm_tf, "peak": tc_peak, "unit": "TFLOP/s", "frac": gemm_tf / tc_peak,
                               "mma_issued_tflops": 3 * gemm_tf, "frac_issued": 3 * gemm_tf / tc_peak,
                               "kernel_ms": gemm_ms, "peak_source": peak_src + " (bf16 dense, burst)"},
            "clocks": clk.summary()}
    if orc is not None:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        t_cpu = cpu_forward_time(orc, G_host, reps=3, warm=1)
        line["cpu_baseline"] = {"value": E / t_cpu, "unit": "edges/s", "cores": cores, "kind": "port",
                                "sample": "median of 3 full config-2 forwards of the oracle (reference-structured "
                                          "PyTorch CPU restatement, not DGL)", "ms_per_step": t_cpu * 1e3}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
