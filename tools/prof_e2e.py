"""Where the end-to-end (host buffers -> logits) time of one config-2 slide goes: H2D, plan build, forward."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from wsi_hgnn_b200.hetero_graph import HeteroGraph

dev = torch.device("cuda", 0)
ours, _ = bench.build_models(False, True)
ours = ours.to(dev)
G_host = bench.make_graph(1)
pinned = HeteroGraph.from_state(G_host.state())
for fr in list(pinned._ndata.values()) + list(pinned._edata.values()):
    for k_ in list(fr):
        fr[k_] = fr[k_].pin_memory()
for ce in list(pinned._edges):
    s_, d_ = pinned._edges[ce]
    pinned._edges[ce] = (s_.pin_memory(), d_.pin_memory())

def sync(): torch.cuda.synchronize()
def t(): sync(); return time.perf_counter()
for it in range(6):
    t0 = t(); g = pinned.to(dev, non_blocking=True); t1 = t()
    plan = g.plan(); t2 = t()
    w = plan.attn_work(); t3 = t()
    with torch.no_grad(): out = ours(g)
    t4 = t(); o = out.cpu(); t5 = t()
    with torch.no_grad(): out = ours(g)
    t6 = t()
    if it >= 3:
        print(f"h2d {1e3*(t1-t0):.3f} ms | plan {1e3*(t2-t1):.3f} | attn_work {1e3*(t3-t2):.3f} | fwd(first on graph) {1e3*(t4-t3):.3f} | d2h {1e3*(t5-t4):.3f} | fwd(again) {1e3*(t6-t5):.3f}")
import cProfile, pstats
g = pinned.to(dev); sync()
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    with torch.no_grad(): ours(g)
sync(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
