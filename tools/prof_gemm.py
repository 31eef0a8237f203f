"""Tiny driver for ncu: tcgen05 typed-linear launches on the config-2 shapes, operands already in operand form.
    python tools/prof_gemm.py [precision]     launches: 3 x K|V|Q (plain epilogue), 3 x a_linear (+ skip mix, FULL epilogue)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wsi_hgnn_b200 import ops
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
N, K, T = 8192, 512, 3
dev = torch.device("cuda", 0)
ptr = [0, N // 3, 2 * (N // 3), N]
with ops.matmul_precision(prec):
    for n_out, full in ((1536, False), (512, True)):
        x = torch.randn(N, K, device=dev); w = torch.randn(T, n_out, K, device=dev) / K ** 0.5; b = torch.randn(T, n_out, device=dev)
        xs, ws = ops.to_operand(x), ops.to_operand(w)
        kw = dict(skip=torch.ones(T, device=dev), res=torch.randn(N, n_out, device=dev), row_gate=torch.ones(N, device=dev)) if full else {}
        for _ in range(3):
            ops.typed_linear_op(xs, ws, b, ptr, n_out, want_op=full, **kw)
torch.cuda.synchronize()
