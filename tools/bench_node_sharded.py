"""BASELINE config 4: one TCGA-COAD-shape slide (100k nodes, F=1024, radius 9 -> 800k edges, T=6), k-NN graph
constructor + HEATNet4 (L=2, D=512, H=4) forward, node-sharded over the GPUs of one box.  One process per GPU:

    python tools/bench_node_sharded.py --nodes 100000                       # 1 GPU (unsharded reference numbers)
    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 tools/bench_node_sharded.py

Checks, in the same run: sharded edge_index == single-rank builder (bit-exact, rank 0 recomputes it when --check),
sharded logits == unsharded CUDA forward (rank 0, --check).  Prints one JSON line on rank 0.
Development / DESIGN.md numbers; bench.py remains the contract benchmark (config 2 forward)."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=100000)
    ap.add_argument("--feat", type=int, default=1024)
    ap.add_argument("--radius", type=int, default=9)
    ap.add_argument("--types", type=int, default=6)
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--hidden", type=int, default=512)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--kv-wire", default="auto", choices=["auto", "fp32", "16"])
    args = ap.parse_args()
    import torch.distributed as dist
    import golden_util
    from wsi_hgnn_b200 import synthetic
    from wsi_hgnn_b200.construct_graph.graph_constructor import construct_graph_arrays
    from wsi_hgnn_b200.hetero_graph import to_heterogeneous
    from wsi_hgnn_b200.models import HEATNet4
    from wsi_hgnn_b200.node_sharded import DistComm, LocalComm, NodeShardedHEAT, knn_edges_sharded

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        comm = DistComm()
    else:
        comm = LocalComm(1).view(0)
    T = args.types
    feats, ntype = synthetic.synth_features(args.nodes, args.feat, T, seed=2, skew=(T == 6))
    feats, ntype = feats.to(dev).contiguous(), ntype.to(dev)

    def sync_time():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return time.perf_counter()

    # ---- edge builder, query rows sharded
    if world > 1:                                                                              # warm-up (one-time set-up costs)
        knn_edges_sharded(feats[:4096].contiguous(), args.radius, comm)
    else:
        construct_graph_arrays(feats[:4096].contiguous(), args.radius)
    t0 = sync_time()
    if world > 1:
        ei, et, sim = knn_edges_sharded(feats, args.radius, comm)
    else:
        ei, et, sim = construct_graph_arrays(feats, args.radius)
    t_knn = sync_time() - t0
    edges_ok = None
    if args.check and world > 1 and rank == 0:
        ei1, et1, sim1 = construct_graph_arrays(feats, args.radius)
        edges_ok = bool(torch.equal(ei, ei1) and torch.equal(et, et1) and torch.equal(sim, sim1))
    G = to_heterogeneous(ei[0], ei[1], ntype, et.to(torch.int64), [str(t) for t in range(T)], ["neg", "pos"],
                         ndata={"feat": feats}, edata={"sim": sim})
    E = G.num_edges()

    torch.manual_seed(611)
    model = HEATNet4(in_dim=args.feat, hidden_dim=args.hidden, out_dim=2, n_layers=args.layers, n_heads=4,
                     node_dict={str(i): i for i in range(T)}, dropuout=0.2)
    golden_util.fill_params(model, 611)
    model = model.to(dev).eval()

    t0 = sync_time()
    sh = NodeShardedHEAT(model, G, comm, kv_wire=args.kv_wire)
    t_plan = sync_time() - t0
    for _ in range(3):
        logits = sh.forward()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync_time()
    for a, b in ev:
        a.record()
        logits = sh.forward()
        b.record()
    sync_time()
    t_fwd = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_fwd, op=dist.ReduceOp.MAX)
        all_logits = [torch.zeros_like(logits) for _ in range(world)]
        dist.all_gather(all_logits, logits)
        same = all(torch.equal(all_logits[0], o) for o in all_logits)
    else:
        same = True
    err = None
    t_single = None
    if args.check and rank == 0:
        with torch.no_grad():
            ref = model(G)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.steps):
                ref = model(G)
            b.record()
            torch.cuda.synchronize()
            t_single = a.elapsed_time(b) / args.steps
        err = float((logits.double() - ref.double()).norm() / ref.double().norm())
    if rank == 0:
        print(json.dumps({"bench": "config4: node-sharded single-slide k-NN builder + HEATNet4 forward", "n_gpus": world,
                          "nodes": args.nodes, "edges": E, "feat": args.feat, "hidden": args.hidden, "layers": args.layers,
                          "knn_pearson_s": t_knn, "plan_s": t_plan, "fwd_ms": float(t_fwd),
                          "fwd_edges_per_s": E / (float(t_fwd) * 1e-3), "rows_per_rank": [sh.bounds[p + 1] - sh.bounds[p] for p in range(world)],
                          "halo_mb_per_layer_per_rank": sh.halo_bytes_per_layer / 1e6, "kv_wire": str(sh.kv_dtype), "logits_identical_on_all_ranks": same,
                          "edge_index_bit_exact_vs_single": edges_ok, "rel_err_vs_unsharded": err,
                          "unsharded_fwd_ms_rank0": t_single, "logits": logits.cpu().tolist()}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
