"""Config-5-shaped data-parallel TRAIN step (BASELINE.json configs[4]): batch of synthetic WSI graphs (2k-20k nodes,
k=8, T=6, F=1024), HEATNet4 D=512 H=4 L=2, CE loss, Adam(lr 1e-5, wd 5e-3), forward + backward + one flat-gradient
all-reduce + optimizer step.  One process per GPU:

    python tools/bench_train.py --batch 32 --steps 5
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_train.py --batch 32

Graphs are built ON THE GPU by the product's own edge builder (k-NN + Pearson kernels).  Prints one JSON line.
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
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--max-nodes", type=int, default=20000)
    ap.add_argument("--cuda-profile", action="store_true",
                    help="cudaProfilerStart/Stop around the timed steps (ncu --profile-from-start off: steady-state launch list)")
    args = ap.parse_args()
    import torch.distributed as dist
    import golden_util
    from wsi_hgnn_b200 import synthetic
    from wsi_hgnn_b200.construct_graph import GraphConstructor
    from wsi_hgnn_b200.hetero_graph import pack
    from wsi_hgnn_b200.models import HEATNet4
    from wsi_hgnn_b200.parallel import FlatModel, flat_train_step
    from wsi_hgnn_b200.sharding import lpt_assign

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    T, F, k = 6, 1024, 8
    g = torch.Generator().manual_seed(1234)
    sizes = torch.randint(2000, args.max_nodes + 1, (args.batch,), generator=g).tolist()
    mine = lpt_assign([s * k for s in sizes], world)[rank]
    t0 = time.perf_counter()
    graphs = []
    for i in mine:
        feats, ntype = synthetic.synth_features(sizes[i], F, T, seed=1000 + i, skew=True)
        het, _, _ = GraphConstructor({"radius": k + 1, "n_node_type": T}, feats, ntype.numpy(), device=dev).construct_graph()
        graphs.append(het)
    G = pack(graphs)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    labels = torch.tensor([i % 2 for i in mine], device=dev)
    model = HEATNet4(in_dim=F, hidden_dim=512, out_dim=2, n_layers=2, n_heads=4, node_dict={str(i): i for i in range(T)},
                     dropuout=0.2)
    golden_util.fill_params(model, 611)
    model = model.to(dev).train()
    red = FlatModel(model)
    G.plan().attn_work()
    for _ in range(args.warmup):
        flat_train_step(model, red, [G], [labels], args.batch, 1e-5, 5e-3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.cuda_profile:
        torch.cuda.profiler.start()
    a.record()
    for _ in range(args.steps):
        loss = flat_train_step(model, red, [G], [labels], args.batch, 1e-5, 5e-3)
    b.record()
    if args.cuda_profile:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) * 1e-3], device=dev, dtype=torch.float64)
    tot = torch.tensor([float(G.num_edges()), float(G.num_nodes()), float(loss)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank == 0:
        sec = float(t) / args.steps
        print(json.dumps({"bench": "config5-shaped HEATNet4 train step (fwd+bwd+allreduce+Adam)", "n_gpus": world,
                          "global_batch": args.batch, "nodes": int(tot[1]), "edges": int(tot[0]), "ms_per_step": sec * 1e3,
                          "graphs_per_s": args.batch / sec, "edges_per_s": float(tot[0]) / sec, "loss_sum": float(tot[2]),
                          "grad_buffer_mb": red.numel * 4 / 1e6, "graph_build_s_rank0": t_build,
                          "peak_mem_gb_rank0": torch.cuda.max_memory_allocated() / 1e9}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
