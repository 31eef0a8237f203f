"""Tiny driver for ncu: edge-attention launches on a DRAM-resident graph (default: config-4 size, 100k nodes, k = 8: K|V
410 MB >> 126 MB L2) or, with --nodes 8192 --k 5, the L2-resident config-2 graph.
    python tools/prof_attn.py [--nodes 100000] [--k 8]"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wsi_hgnn_b200 import ops, synthetic
ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=100000)
ap.add_argument("--k", type=int, default=8)
args = ap.parse_args()
N, T, D, H = args.nodes, 6 if args.nodes > 20000 else 3, 512, 4
dev = torch.device("cuda", 0)
G = synthetic.device_slide_graph(N, 64, T, args.k, seed=2, device=dev, skew=T == 6)
plan = G.plan()
kvq = torch.randn(N, 3 * D, device=dev)
ew, eb = torch.ones(1, device=dev), torch.zeros(1, device=dev)
work = plan.attn_work()
for _ in range(4):
    kvq.add_(0.0)
    ops.hetero_attn_work(kvq[:, :D], kvq[:, D:2 * D], kvq[:, 2 * D:], work, plan.e_src, plan.e_sim,
                         plan.e_rel, plan.node_inv_r, ew, eb, D, H, op_out=True)
torch.cuda.synchronize()
E = G.num_edges()
deg = (plan.rowptr[1:] - plan.rowptr[:-1]).float()
print("nodes %d edges %d in-degree: mean %.2f max %d zero %d; algorithmic bytes per launch %d" %
      (N, E, deg.mean().item(), int(deg.max()), int((deg == 0).sum()), E * (2 * D * 4 + 8) + N * (2 * D * 4 + 4)))
