"""Tiny driver for ncu: the batch-scale tcgen05 launches (a packed training micro-batch, 163 840 rows, 6 skewed types):
    python tools/prof_wgrad.py      launches: K|V|Q forward GEMM bf16x3 + fp16, typed_wgrad K|V|Q and a_linear (+ their reduce)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wsi_hgnn_b200 import ops, synthetic
N, T = 163840, 6
dev = torch.device("cuda", 0)
cnt = [int(N * f) for f in synthetic.TYPE_SKEW6]
cnt[0] += N - sum(cnt)
tp = [0]
for c in cnt:
    tp.append(tp[-1] + c)
x = torch.randn(N, 512, device=dev)
for n_out in (1536, 512):
    w = torch.randn(T, n_out, 512, device=dev) / 512 ** 0.5
    b = torch.randn(T, n_out, device=dev)
    dy = torch.randn(N, n_out, device=dev)
    for prec in ("bf16x3", "fp16"):
        with ops.matmul_precision(prec):
            xs, ws = ops.to_operand(x), ops.to_operand(w)
            for _ in range(2):
                ops.typed_linear_op(xs, ws, b, tp, n_out)
    xs, ds = ops.to_operand(x, ops.OPF_BF16X3), ops.to_operand(dy, ops.OPF_BF16X3)
    for _ in range(2):
        ops.typed_wgrad(ds, xs, tp)
torch.cuda.synchronize()
