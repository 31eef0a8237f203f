"""Diagnostic: which part of bench.py's setup slows slide_io.stream_forward down (development only)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from wsi_hgnn_b200.graphed import GraphedForward
from wsi_hgnn_b200.slide_io import FlatSlide, stream_forward

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
ours, _ = bench.build_models(False, True)
ours = ours.to(dev)
G_host = bench.make_graph(1)
slides = [FlatSlide.from_graph(G_host if i == 0 else bench.make_graph(101 + i), pin=True) for i in range(4)]
many = [slides[i % 4] for i in range(24)]
print("pinned:", [s.blob.is_pinned() for s in slides])

def run(tag):
    list(stream_forward(ours, many[:6], dev))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    list(stream_forward(ours, many, dev))
    torch.cuda.synchronize(); t1 = time.perf_counter()
    s = slides[1]
    torch.cuda.synchronize(); t2 = time.perf_counter()
    b = s.blob[:s.header["nbytes"]].to(dev, non_blocking=True)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"{tag}: stream {1e3*(t1-t0)/len(many):.3f} ms/slide | one H2D {1e3*(t3-t2):.3f} ms", flush=True)

run("fresh")
G = G_host.to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
run("after G.to + flush alloc")
gf = GraphedForward(ours, G, warmup=3)
for _ in range(5):
    gf()
torch.cuda.synchronize()
run("after GraphedForward")
with bench.ClockSampler(0) as clk:
    for _ in range(30):
        flush.zero_(); gf()
    torch.cuda.synchronize()
print(clk.summary())
run("after ClockSampler")
time.sleep(3)
run("3 s later")
