"""Tiny driver for ncu: a few tcgen05 typed-linear launches on the config-2 K|V|Q shape."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wsi_hgnn_b200 import ops
N, K, n_out, T = 8192, 512, int(sys.argv[1]) if len(sys.argv) > 1 else 1536, 3
dev = torch.device("cuda", 0)
x = torch.randn(N, K, device=dev); w = torch.randn(T, n_out, K, device=dev) / K ** 0.5; b = torch.randn(T, n_out, device=dev)
ptr = [0, N // 3, 2 * (N // 3), N]
out = torch.empty(N, n_out, device=dev)
for _ in range(5):
    ops.typed_linear(x, w, b, ptr, impl=ops.IMPL_TC, out=out)
torch.cuda.synchronize()
