"""In-situ time of every kernel of the config-2 HEATNet4 forward chain (operands in operand form, the launches of
wsi_heat_forward issued one by one from Python with a CUDA event between them; L2 is flushed once per forward, so every
kernel sees the cache state the previous kernels leave - unlike the per-kernel cold numbers of tools/bench_kernels.py).

    python tools/forward_breakdown.py [--precision fp16] [--reps 20]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    sys.argv = [sys.argv[0]]
    import bench
    from wsi_hgnn_b200 import ops
    from wsi_hgnn_b200.models.heat import _graph_type_order, packed_features, readout_scale
    from wsi_hgnn_b200.models._packing import stack_linears
    dev = torch.device("cuda", 0)
    ops.set_matmul_precision(args.precision)
    ours, _ = bench.build_models(False, True)
    ours = ours.to(dev)
    G = bench.make_graph(1).to(dev)
    plan = G.plan()
    order = _graph_type_order(plan, ours.node_dict)
    D, H = bench.CFG["hidden"], bench.CFG["heads"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    with torch.no_grad():
        feat = packed_features(G, plan, None)
        w_in, b_in = stack_linears(ours.adapt_ws, order)
        w_in_s = ops.to_operand(w_in)
        packs = [(l._packed(order), l._packed_split(order)) for l in ours.gcs]
        work = plan.attn_work()
        names = list(plan.ntypes)
        M, c, b_total = ours._affine_maps(names, True)
        scale = readout_scale(plan, False)
        tpc = plan.type_ptr_c()
        # every stage is captured into its own CUDA graph (inputs = the outputs the previous stage's capture produced, which
        # stay alive in that graph's pool), so that the timed sequence costs the host one cheap replay per stage and the
        # GPU timeline stays dense
        state = {}
        stage_fns = []

        def add(name, fn):
            stage_fns.append((name, fn))

        # (every stage writes its OWN key: the eager warm-up calls before a capture must not replace a tensor another
        #  stage's captured graph reads)
        add("to_operand(feat)", lambda: state.__setitem__("fs", ops.to_operand(feat)))
        add("adapt_ws GEMM", lambda: state.__setitem__("x0", ops.typed_linear_op(state["fs"], w_in_s, b_in, plan.type_ptr, D, want_op=True, type_ptr_c=tpc)))
        for i, layer in enumerate(ours.gcs):
            (w_kvq, b_kvq, wa, ba, skip, _), (w_kvq_s, wa_s) = packs[i]
            add(f"L{i} K|V|Q GEMM", lambda w=w_kvq_s, b=b_kvq, i=i: state.__setitem__(f"kvq{i}", ops.typed_linear_op(state[f"x{i}"][1], w, b, plan.type_ptr, 3 * D, type_ptr_c=tpc)[0]))
            add(f"L{i} edge attention", lambda layer=layer, i=i: state.__setitem__(f"agg{i}", ops.hetero_attn_work(
                state[f"kvq{i}"][:, :D], state[f"kvq{i}"][:, D:2 * D], state[f"kvq{i}"][:, 2 * D:], work, plan.e_src, plan.e_sim, plan.e_rel,
                plan.node_inv_r, layer.e_linear.weight, layer.e_linear.bias, D, H, op_out=True)))
            add(f"L{i} a_linear GEMM + skip mix", lambda w=wa_s, b=ba, sk=skip, i=i, last=(i + 1 == len(ours.gcs)): state.__setitem__(
                f"x{i + 1}", ops.typed_linear_op(state[f"agg{i}"], w, b, plan.type_ptr, D, skip=sk, res=state[f"x{i}"][0], row_gate=plan.node_inv_r,
                                              want_op=not last, type_ptr_c=tpc)))
        add("readout (pool + affine heads)", lambda: state.__setitem__("out", ops.segment_pool_affine(
            state[f"x{len(ours.gcs)}"][0], plan.seg_ptr, len(names), plan.B, ours.graph_pooling_type, M, c, b_total, scale)))
        graphs = []
        for name, fn in stage_fns:
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            graphs.append((name, g))
        torch.cuda.synchronize()
        acc = {}
        for _ in range(args.reps):
            evs = []
            flush.zero_()
            for name, g in graphs:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                evs.append((name, e))
                g.replay()
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            evs.append((None, e))
            torch.cuda.synchronize()
            for (n, a), (_, b) in zip(evs, evs[1:]):
                acc.setdefault(n, []).append(a.elapsed_time(b) * 1e3)
        total = 0.0
        for n, ts in acc.items():
            ts.sort()
            med = ts[len(ts) // 2]
            total += med
            print(json.dumps({"stage": n, "us_median": round(med, 2), "us_min": round(ts[0], 2)}), flush=True)
        print(json.dumps({"stage": "sum of medians (one CUDA graph per stage, back to back)", "us": round(total, 1), "precision": args.precision}))


if __name__ == "__main__":
    main()
