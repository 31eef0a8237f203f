"""Where the end-to-end (pinned host blob -> logits) time of config-2 slides goes: stage timings of one slide and a
cProfile of slide_io.stream_forward (the host side is what bounds the streamed path)."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from wsi_hgnn_b200.slide_io import FlatSlide, stream_forward  # noqa: E402

dev = torch.device("cuda", 0)
ours, _ = bench.build_models(False, True)
ours = ours.to(dev)
slides = [FlatSlide.from_graph(bench.make_graph(1 + i), pin=True) for i in range(4)]


def t():
    torch.cuda.synchronize()
    return time.perf_counter()


for it in range(6):
    s = slides[it % 4]
    t0 = t(); blob = s.blob[:s.header["nbytes"]].to(dev, non_blocking=True); t1 = t()
    g = s.graph_on(blob); t1b = t()
    plan = g.plan(); t2 = t()
    ours.prepare_plan(plan); t3 = t()
    with torch.no_grad():
        out = ours(g)
    t4 = t(); o = out.cpu(); t5 = t()
    if it >= 3:
        print(f"h2d {1e3*(t1-t0):.3f} ms | graph views {1e3*(t1b-t1):.3f} | plan {1e3*(t2-t1b):.3f} | work list {1e3*(t3-t2):.3f} "
              f"| fwd {1e3*(t4-t3):.3f} | d2h {1e3*(t5-t4):.3f}")
many = [slides[i % 4] for i in range(40)]
list(stream_forward(ours, many[:8], dev))
t0 = t(); outs = list(stream_forward(ours, many, dev)); t1 = t()
print(f"stream_forward: {1e3*(t1-t0)/len(many):.3f} ms / slide")
pr = cProfile.Profile(); pr.enable()
outs = list(stream_forward(ours, many, dev))
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
