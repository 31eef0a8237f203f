"""CUDA-graph replay of a model forward on a fixed slide graph.

One forward is ~20 short kernels over a few thousand edges each, so at single-slide sizes it is launch-bound when
issued from Python; capturing it once and replaying removes the host from the loop.  New node features are fed by
copying into the graph's packed feature buffer (``set_features``); the logits tensor is static.
"""
import torch

from . import _lib
from .models.heat import packed_features


class GraphedForward:
    def __init__(self, model, G, warmup: int = 3):
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedForward needs a CUDA device")
        self.model = model.eval()
        self.G = G
        self.plan = G.plan()
        self.feat = packed_features(G, self.plan, None)          # static input buffer [N, F]
        lib = _lib.load()
        side = torch.cuda.Stream(device=G.device)
        side.wait_stream(torch.cuda.current_stream(G.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):                      # warms the weight-pack and plan caches
                model(G)
        torch.cuda.current_stream(G.device).wait_stream(side)
        torch.cuda.synchronize(G.device)
        self.graph = torch.cuda.CUDAGraph()
        before = lib.wsi_launch_count()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = model(G)
        self.kernels_per_replay = int(lib.wsi_launch_count() - before)

    def set_features(self, feat: torch.Tensor, non_blocking: bool = True):
        """Copy new packed [N, F] features (host or device) into the captured input buffer."""
        self.feat.copy_(feat, non_blocking=non_blocking)

    def __call__(self) -> torch.Tensor:
        self.graph.replay()
        return self.out
