"""Where does the streamed end-to-end path spend its time?  wsi_stream_forward over 32 distinct config-2 slides with the
development knob stream_debug: copies only / plan + forward only / everything.    python tools/diag_stream.py"""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.argv = [sys.argv[0]]
import bench
from wsi_hgnn_b200 import ops, synthetic
from wsi_hgnn_b200.slide_io import FlatSlide, stream_forward
dev = torch.device("cuda", 0)
ours, _ = bench.build_models(False, True)
ours = ours.to(dev)
C = bench.CFG
for feat_dtype in ("fp16", "fp32"):
    slides = [FlatSlide.from_graph(synthetic.device_slide_graph(C["nodes"], C["in_dim"], C["node_types"], C["k"], seed=900 + i, device=dev).to("cpu"),
                                   pin=True, feat_dtype=feat_dtype) for i in range(32)]
    list(stream_forward(ours, slides[:6], dev))
    for dbg, tag in ((0, "everything"), (2, "copies only"), (1, "plan + forward only (blobs copied once)"), (3, "host loop only")):
        ops.dev_set("stream_debug", dbg)
        ts = []
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            list(stream_forward(ours, slides if not (dbg & 1) else [slides[0]] * len(slides), dev))    # (stale blobs need one slide)
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) / len(slides) * 1e3)
        print(json.dumps({"feat": feat_dtype, "mode": tag, "ms_per_slide": min(ts), "all": ts}), flush=True)
    ops.dev_set("stream_debug", 0)
    for loop in ("python",):
        os.environ["WSI_STREAM_LOOP"] = loop
        ts = []
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            list(stream_forward(ours, slides, dev))
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) / len(slides) * 1e3)
        print(json.dumps({"feat": feat_dtype, "mode": "python-issued pipeline", "ms_per_slide": min(ts), "all": ts}), flush=True)
        os.environ.pop("WSI_STREAM_LOOP")
