"""Development: first contact with the 4-CTA-cluster (A-multicast) variant of the tcgen05 GEMM - run under `timeout`."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wsi_hgnn_b200 import ops

dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
os.environ["WSI_TC_VERBOSE"] = "1"
for (N, K, n_out, T) in ((8192, 512, 1536, 3), (8192, 512, 512, 3), (8192, 1024, 512, 3), (3001, 128, 512, 2), (700, 64, 1024, 1)):
    ptr = [0] + sorted(torch.randint(1, N, (T - 1,), generator=g).tolist()) + [N]
    x = torch.randn(N, K, generator=g).to(dev)
    w = (torch.randn(T, n_out, K, generator=g) / K ** 0.5).to(dev)
    b = torch.randn(T, n_out, generator=g).to(dev)
    xs, ws = ops.split_bf16(x), ops.split_bf16(w)
    res = torch.randn(N, n_out, generator=g).to(dev)
    kw = dict(skip=torch.ones(T, device=dev), res=res, row_gate=torch.ones(N, device=dev))
    for full in (False, True):
        os.environ["WSI_TC_CL"] = "2"
        y2, s2 = ops.typed_linear_split(xs, ws, b, ptr, n_out, want_split=True, **(kw if full else {}))
        torch.cuda.synchronize()
        os.environ["WSI_TC_CL"] = "4"
        t0 = time.time()
        y4, s4 = ops.typed_linear_split(xs, ws, b, ptr, n_out, want_split=True, **(kw if full else {}))
        torch.cuda.synchronize()
        ok = torch.equal(y2, y4) and torch.equal(s2, s4)
        ref = torch.cat([x[ptr[t]:ptr[t + 1]].double() @ w[t].double().T + b[t].double() for t in range(T)])
        print(f"N={N} K={K} n_out={n_out} T={T} full={full}: CL4 == CL2 bit-exact: {ok}; max |diff| {float((y2 - y4).abs().max()):.3e}; "
              f"{'(plain) rel err vs fp64 %.2e' % float((y4.double() - ref).norm() / ref.norm()) if not full else ''} [{time.time() - t0:.2f}s]", flush=True)
