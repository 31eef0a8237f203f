"""bench.py - HEATNet4 forward throughput (edges/s) on the BASELINE.json config-2 workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one HEATNet4 forward (input projection + 3 HEAT layers + typed readout + heads) over one synthetic
TCGA-BRCA-shape slide graph (8192 patch nodes, 3 node types, 40 960 edges, F=1024, D=512, H=4, fp32).
  value     : edges/s with the graph resident in HBM (CSR pre-built), CUDA-graph replay, CUDA-event timed,
              L2 flushed between steps
  e2e       : the same metric through the public API from pinned HOST buffers (slide_io.stream_forward): per slide
              one H2D copy of features + edge arrays, CSR build, forward, logits D2H - all inside the timed region;
              every timed slide is a DISTINCT slide whose host-side plan has not been computed before
  roofline  : the edge-attention kernel (gather K/V by source, per-relation softmax, scatter to dst), HBM bound
  roofline_dense : the tcgen05 typed GEMM (K|V|Q projection), tensor-pipe bound
  train_step: BASELINE config 5 - the 256-graph batch (2k-20k nodes each) LPT-sharded over the N ranks, full step =
              forward + backward + bucketed gradient all-reduce (NCCL) + Adam; STRONG scaling (same global batch for
              every N), per-rank phase times
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
            "config": dict(CFG),
            "notes": {"arm": "reference-structured PyTorch CPU restatement (oracle/), not DGL"},
            "cpu_baseline": {"value": v, "unit": "edges/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} full config-2 forwards"},
            "e2e": {"value": v, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_train_step(args, rank, world, dev):
    """BASELINE config 5: the 256-graph batch, LPT-sharded over the ranks, fwd + bwd + bucketed all-reduce + Adam."""
    import torch.distributed as dist
    import golden_util
    from wsi_hgnn_b200 import synthetic
    from wsi_hgnn_b200.hetero_graph import pack
    from wsi_hgnn_b200.models import HEATNet4
    from wsi_hgnn_b200.parallel import FlatModel, flat_train_step
    from wsi_hgnn_b200.sharding import lpt_assign
    B, T, F, k = args.train_graphs, 6, 1024, 8
    g = torch.Generator().manual_seed(1234)
    sizes = torch.randint(2000, 20001, (B,), generator=g).tolist()
    mine = lpt_assign([s * k for s in sizes], world)[rank]
    t0 = time.perf_counter()
    # micro-batches of <= ~160k nodes bound the activation memory (gradient accumulation); plans (CSR + work lists) are
    # built once, outside the timed steps, like the resident graph of `value`
    packs, labels, cur, cur_n = [], [], [], 0
    def flush_pack():
        nonlocal cur, cur_n
        if cur:
            G = pack([c[1] for c in cur])
            G.plan().attn_work()
            G.plan().rows_by_degree()
            packs.append(G)
            labels.append(torch.tensor([c[0] % 2 for c in cur], device=dev))
            cur, cur_n = [], 0
    for i in mine:
        if cur_n + sizes[i] > args.train_pack_nodes:
            flush_pack()
        cur.append((i, synthetic.device_slide_graph(sizes[i], F, T, k, seed=1000 + i, device=dev, skew=True)))
        cur_n += sizes[i]
    flush_pack()
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    n_nodes = sum(G.num_nodes() for G in packs)
    n_edges = sum(G.num_edges() for G in packs)
    model = HEATNet4(in_dim=F, hidden_dim=512, out_dim=2, n_layers=2, n_heads=4, node_dict={str(i): i for i in range(T)},
                     dropuout=0.2)
    golden_util.fill_params(model, 611)
    model = model.to(dev).train()
    flat = FlatModel(model, bucket_mb=16.0)
    lr, wd = 1e-5, 5e-3                                    # configs/BRCA/HEAT4_kimia_classification_v2.yml
    for _ in range(max(1, args.train_warmup)):
        flat_train_step(model, flat, packs, labels, B, lr, wd)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    events = {}
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.train_steps):
        loss = flat_train_step(model, flat, packs, labels, B, lr, wd, events=events)
    b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = a.elapsed_time(b) / args.train_steps
    phase = {k_: sum(x.elapsed_time(y) for x, y in v) / args.train_steps for k_, v in events.items()}
    mine_stats = torch.tensor([ms, phase.get("fwd", 0.0), phase.get("bwd", 0.0), phase.get("comm", 0.0), phase.get("opt", 0.0),
                               float(n_nodes), float(n_edges), float(len(mine)), float(loss), t_build,
                               torch.cuda.max_memory_allocated() / 1e9], device=dev, dtype=torch.float64)
    allr = [torch.zeros_like(mine_stats) for _ in range(world)]
    if world > 1:
        dist.all_gather(allr, mine_stats)
    else:
        allr = [mine_stats]
    if rank != 0:
        return None
    rows = [r.tolist() for r in allr]
    t_max = max(r[0] for r in rows)
    tot_edges = sum(r[6] for r in rows)
    return {"workload": "config5: batch of 256 synthetic WSI graphs (2k-20k nodes, k=8, T=6, F=1024), HEATNet4 D=512 H=4 L=2 "
                        "dropout 0.2, CE loss, Adam(lr 1e-5, wd 5e-3): forward + backward + gradient all-reduce + optimizer",
            "global_batch": B, "n_gpus": world, "scaling": "strong", "steps": args.train_steps, "warmup": max(1, args.train_warmup),
            "ms_per_step": t_max, "graphs_per_s": B / (t_max * 1e-3), "edges_per_s": tot_edges / (t_max * 1e-3),
            "nodes": int(sum(r[5] for r in rows)), "edges": int(tot_edges), "loss_sum": sum(r[8] for r in rows),
            "grad_buffer_mb": flat.numel * 4 / 1e6, "buckets": len(flat.buckets), "sharding": "greedy LPT on edge count",
            "matmul_precision": {"forward": "bf16x3 (3-term split)", "gradients": "bf16x3"},
            "per_rank": [{"rank": i, "ms_per_step": r[0], "fwd_ms": r[1], "bwd_ms": r[2], "allreduce_exposed_ms": r[3],
                          "optimizer_ms": r[4], "nodes": int(r[5]), "edges": int(r[6]), "graphs": int(r[7]),
                          "graph_build_s": r[9], "peak_mem_gb": r[10]} for i, r in enumerate(rows)],
            "note": "time = CUDA events around the timed steps, max over ranks; allreduce_exposed_ms = wait for the bucketed "
                    "all-reduces after the last backward (the buckets are issued during it); graphs are built on the GPU "
                    "by the product's own k-NN + Pearson kernels, plans outside the timed region"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the config-5 train_step record")
    ap.add_argument("--train-graphs", type=int, default=256)
    ap.add_argument("--train-steps", type=int, default=3)
    ap.add_argument("--train-warmup", type=int, default=1)
    ap.add_argument("--train-pack-nodes", type=int, default=160000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from wsi_hgnn_b200 import _lib, ops, synthetic
    from wsi_hgnn_b200.graphed import GraphedForward

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    # torchrun pins OMP_NUM_THREADS=1; the synthetic slides are generated on the host (exact fp64 k-NN), so give every
    # rank its share of the cores for that set-up phase
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // max(world, 1)))
    torch.cuda.set_device(local)
    from wsi_hgnn_b200.sharding import bind_to_gpu_numa
    numa = bind_to_gpu_numa(local)            # before any pinned allocation: NUMA-local staging buffers
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
    precision = ops.get_matmul_precision()

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
    # public API: slide_io.FlatSlide (one pinned blob per slide) + slide_io.stream_forward (the H2D copy of slide i+2 and
    # the CSR / work-list build of slide i+1 overlap the forward of slide i; logits come back through pinned memory).
    # Every timed slide is a DISTINCT slide: copied host -> device, planned and run; nothing is reused between slides
    # (the host-side plan head of each FlatSlide is cleared before the timed pass).
    from wsi_hgnn_b200.slide_io import FlatSlide, stream_forward
    n_e2e = max(8, min(args.steps, 32))

    def make_slides(feat_dtype, n):
        out = []
        for i in range(n):
            Gs = synthetic.device_slide_graph(CFG["nodes"], CFG["in_dim"], CFG["node_types"], CFG["k"],
                                              seed=5000 + 97 * rank + i, device=dev)
            out.append(FlatSlide.from_graph(Gs.to("cpu"), pin=True, feat_dtype=feat_dtype))
        return out

    def time_stream(slides):
        list(stream_forward(ours, slides[:4], dev))                 # warm-up: buffers, workspaces, streams
        for s_ in slides:
            s_._head = None                                         # no host plan survives from the warm-up
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        outs = list(stream_forward(ours, slides, dev))
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        ed = torch.tensor([float(sum(s_.num_edges() for s_ in slides))], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dist.all_reduce(ed, op=dist.ReduceOp.SUM)
        return float(ed) / float(dt), float(dt) / len(slides) * 1e3, outs

    slides16 = make_slides("fp16", n_e2e)
    e2e_value, e2e_ms, outs = time_stream(slides16)
    h2d = sum(s_.header["nbytes"] for s_ in slides16) // n_e2e
    d2h = outs[0].numel() * 4
    # stage times of one slide, each alone: the H2D copy of its blob, and (above) the device forward
    dbuf = torch.empty(slides16[0].header["nbytes"], dtype=torch.uint8, device=dev)
    ca, cb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dbuf.copy_(slides16[0].blob[:dbuf.numel()], non_blocking=True)
    torch.cuda.synchronize()
    ca.record()
    for s_ in slides16[:8]:
        dbuf.copy_(s_.blob[:dbuf.numel()], non_blocking=True)
    cb.record()
    torch.cuda.synchronize()
    h2d_ms = ca.elapsed_time(cb) / 8
    # the same copies issued by ALL ranks at once: the aggregate host -> device bandwidth of the box bounds e2e at N > 1
    h2d_all_ms = h2d_ms
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        ca.record()
        for s_ in slides16[:8]:
            dbuf.copy_(s_.blob[:dbuf.numel()], non_blocking=True)
        cb.record()
        torch.cuda.synchronize()
        tt = torch.tensor([ca.elapsed_time(cb) / 8], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        h2d_all_ms = float(tt)
    # the same with fp32 feature blobs (twice the copy), fewer slides
    slides32 = make_slides("fp32", 8)
    e2e32_value, e2e32_ms, _ = time_stream(slides32)
    h2d32 = slides32[0].header["nbytes"]
    # one slide at a time with a host sync per slide (what the reference's evaluation loop does)
    for s_ in slides16[:3]:
        with torch.no_grad():
            ours(s_.to_graph(dev, non_blocking=True)).cpu()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s_ in slides16[:8]:
        with torch.no_grad():
            ours(s_.to_graph(dev, non_blocking=True)).cpu()
    e2e_sync_ms = (time.perf_counter() - t0) / 8 * 1e3
    del slides32, dbuf

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

        def graphed(fn):
            """capture one call of fn: the replay costs no Python / ctypes time between the timing events"""
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):
                fn()
            return g_

        g_attn = graphed(lambda: ops.hetero_attn_work(*attn_args, op_out=True))
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
        # dense: fused K|V|Q typed GEMM on operands in operand form (as inside the forward)
        g_gemm = graphed(lambda: ops.typed_linear_op(xs, w_kvq_s, b_kvq, plan.type_ptr, 3 * D))
        gev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in gev:
            flush.zero_()
            a.record()
            g_gemm.replay()
            b.record()
        torch.cuda.synchronize()
        gemm_ms = sorted(a.elapsed_time(b) for a, b in gev)[reps // 2]
        # the same kernel, and the weight-gradient kernel, at the size of one packed training micro-batch (config 5:
        # 163 840 rows, 6 skewed node types) - where a launch has ~50 tiles per CTA pair instead of 3
        from wsi_hgnn_b200 import synthetic as _syn
        Nb, Tb = 163840, 6
        cnt = [int(Nb * f) for f in _syn.TYPE_SKEW6]
        cnt[0] += Nb - sum(cnt)
        tpb = [0]
        for c in cnt:
            tpb.append(tpb[-1] + c)
        xb = torch.randn(Nb, D, device=dev)
        wb = torch.randn(Tb, 3 * D, D, device=dev) / D ** 0.5
        bb = torch.randn(Tb, 3 * D, device=dev)
        dyb = torch.randn(Nb, 3 * D, device=dev)
        yb = torch.empty(Nb, 3 * D, device=dev)
        xb3, wb3, dy3 = ops.to_operand(xb, ops.OPF_BF16X3), ops.to_operand(wb, ops.OPF_BF16X3), ops.to_operand(dyb, ops.OPF_BF16X3)
        del dyb
        batch_ms = {}
        for tag, fn in (("fwd_bf16x3", lambda: ops.typed_linear_op(xb3, wb3, bb, tpb, 3 * D, opf=ops.OPF_BF16X3, out=yb)),
                        ("wgrad_bf16x3", lambda: ops.typed_wgrad(dy3, xb3, tpb))):
            for _ in range(2):
                fn()
            ts = []
            for _ in range(5):
                flush.zero_()
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record()
                fn()
                b_.record()
                torch.cuda.synchronize()
                ts.append(a_.elapsed_time(b_))
            batch_ms[tag] = sorted(ts)[2]
        del xb, wb, bb, yb, xb3, wb3, dy3
        batch_flops = 2.0 * Nb * D * 3 * D
    attn_bytes = E * (2 * D * 4 + 8) + N * (2 * D * 4 + 4)          # SURVEY.md §8(d) per-layer edge-phase bytes
    achieved = attn_bytes / (attn_ms * 1e-3) / 1e9
    gemm_flops = 2.0 * N * D * 3 * D
    gemm_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12
    issued = 3 if precision == "bf16x3" else 1

    launches_per_fwd = gf.kernels_per_replay
    del gf, g_attn, g_gemm
    train = None
    if not args.no_train:
        torch.cuda.empty_cache()
        train = run_train_step(args, rank, world, dev)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    line = {"metric": METRIC, "value": value, "unit": "edges/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_max / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(CFG),
            "notes": {"l2": "flushed between timed iterations (256 MB write)", "launch": "CUDA-graph replay",
                      "eager_ms_per_step": eager_ms, "parallelism": f"dp{world} (independent slides per rank)",
                      "matmul_precision": f"{precision}: fp32 storage, tensor-core operands rounded to " +
                                          {"fp16": "fp16 (11-bit significand, one pass)", "bf16": "bf16 (one pass)",
                                           "bf16x3": "a 3-term bf16 split"}[precision] + ", fp32 accumulate; parity "
                                          "margin measured in profiles/r2_precision_study.json"},
            "e2e": {"value": e2e_value, "unit": "edges/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "steps": n_e2e, "distinct_slides": n_e2e,
                    "h2d_alone_ms": h2d_ms, "h2d_gb_s": h2d / (h2d_ms * 1e-3) / 1e9, "device_forward_ms": t_max / args.steps * 1e3,
                    "h2d_all_ranks_concurrent_ms": h2d_all_ms,
                    "h2d_all_ranks_aggregate_gb_s": world * h2d / (h2d_all_ms * 1e-3) / 1e9,
                    "rank0_numa_binding": numa,
                    "fp32_feature_blobs": {"value": e2e32_value, "ms_per_step": e2e32_ms, "h2d_bytes_per_step": h2d32, "steps": 8},
                    "sync_ms_per_step": e2e_sync_ms,
                    "note": "slide_io.stream_forward over pinned FlatSlide blobs, every slide distinct and never seen before: per "
                            "slide ONE H2D copy (fp16 features - the operand the fp16 GEMM forms anyway, bit-identical logits - "
                            "+ edges + sim), CSR + work-list build (wsi_slide_plan) and forward (wsi_slide_run), logits D2H; "
                            "three slides in flight (copy | plan | forward).  sync_ms_per_step = slides one at a time "
                            "through model(G) with a host sync per slide"},
            "gpu_launches": launches_per_fwd * args.steps,
            "roofline": {"kernel": "attn_fwd_vec_kernel, one layer (wsi_hetero_attn_work_fwd: 16-byte register gathers of "
                                   "K|V rows by source, per-relation online softmax, fused merge of hub-row chunks, fp16 "
                                   "operand-form write to dst)", "bound": "hbm",
                         "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": ncu_traffic("attn"), "peak_source": peak_src, "algorithmic_bytes": attn_bytes,
                         "kernel_ms": attn_ms, "note": "algorithmic bytes = SURVEY 8(d) edge-phase bytes (every gathered "
                         "K/V row counted); K|V (33.5 MB) is L2 resident at this size, so DRAM traffic is far below it: an "
                         "L2 -> SM figure.  DRAM-resident sizes: profiles/r2_attn_hbm.txt"},
            "roofline_dense": {"kernel": f"typed_linear_tc_kernel K|V|Q [8192x512]x[512x1536] on tcgen05, {precision} operands, "
                                         "fp32 accumulate in TMEM, fp32 output through TMA stores", "bound": "tensor",
                               "achieved": gemm_tf, "peak": tc_peak, "unit": "TFLOP/s", "frac": gemm_tf / tc_peak,
                               "mma_issued_tflops": issued * gemm_tf, "frac_issued": issued * gemm_tf / tc_peak,
                               "kernel_ms": gemm_ms, "peak_source": peak_src + " (bf16 dense, burst)"},
            "roofline_dense_batch": {
                "what": "the tcgen05 GEMMs of the training path at the size of one packed micro-batch of config 5 "
                        f"([{Nb}, {D}] x [{Tb}, {3 * D}, {D}], 6 skewed node types), 3-term bf16 split (fp32-grade), measured live",
                "bound": "tensor", "unit": "TFLOP/s", "peak": tc_peak, "peak_source": peak_src + " (bf16 dense, burst)",
                "forward": {"kernel": "typed_linear_tc_kernel<TERMS=3>", "kernel_ms": batch_ms["fwd_bf16x3"],
                            "achieved": batch_flops / (batch_ms["fwd_bf16x3"] * 1e-3) / 1e12,
                            "frac_issued": 3 * batch_flops / (batch_ms["fwd_bf16x3"] * 1e-3) / 1e12 / tc_peak},
                "wgrad": {"kernel": "typed_wgrad_tc_kernel (MN-major operands) + wgrad_reduce_kernel", "kernel_ms": batch_ms["wgrad_bf16x3"],
                          "achieved": batch_flops / (batch_ms["wgrad_bf16x3"] * 1e-3) / 1e12,
                          "frac_issued": 3 * batch_flops / (batch_ms["wgrad_bf16x3"] * 1e-3) / 1e12 / tc_peak},
                "ncu": "profiles/r2_batch_scale_tensor_pipe.txt (sm__pipe_tensor_cycles_active 92 % / 78 %)"},
            "clocks": clk.summary()}
    if train is not None:
        line["train_step"] = train
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
