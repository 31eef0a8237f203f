"""Tiny driver for ncu: the config-4 edge builder (100k nodes, F=1024, radius 9) once, after one warm-up."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wsi_hgnn_b200 import synthetic
from wsi_hgnn_b200.construct_graph.graph_constructor import construct_graph_arrays
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
feats, _ = synthetic.synth_features(n, 1024, 6, seed=7, skew=True, device="cuda")
construct_graph_arrays(feats, 9)
torch.cuda.synchronize()
torch.cuda.profiler.start()
t0 = time.perf_counter()
ei, et, sim = construct_graph_arrays(feats, 9)
torch.cuda.synchronize()
print("builder_s", time.perf_counter() - t0, "edges", ei.shape)
torch.cuda.profiler.stop()
