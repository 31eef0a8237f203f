"""Tiny driver for ncu: a few edge-attention launches on the config-2 graph."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wsi_hgnn_b200 import ops, synthetic
N, T, D, H = 8192, 3, 512, 4
dev = torch.device("cuda", 0)
G = synthetic.synth_slide_graph(N, 64, T, 5, seed=1).to(dev)
plan = G.plan()
kvq = torch.randn(N, 3 * D, device=dev)
ew, eb = torch.ones(1, device=dev), torch.zeros(1, device=dev)
use_perm = ops.head_perm(D, H) is not None
work = plan.attn_work()
for _ in range(5):
    kvq.add_(0.0)
    ops.hetero_attn_work(kvq[:, :D], kvq[:, D:2 * D], kvq[:, 2 * D:], work, plan.e_src, plan.e_sim,
                         plan.e_rel, plan.node_inv_r, ew, eb, D, H)
torch.cuda.synchronize()
deg = (plan.rowptr[1:] - plan.rowptr[:-1]).float()
print("in-degree: mean %.2f max %d zero %d" % (deg.mean().item(), int(deg.max()), int((deg == 0).sum())))
