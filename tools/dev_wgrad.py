"""Development check of the tcgen05 weight-gradient kernel against fp64 torch (and timing vs the cuBLAS route)."""
import os, sys, json, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wsi_hgnn_b200 import ops


def ref(dy, x, tp):
    return torch.stack([dy[tp[t]:tp[t + 1]].double().t() @ x[tp[t]:tp[t + 1]].double() for t in range(len(tp) - 1)])


def cublas_route(ds, xs, tp):
    from wsi_hgnn_b200.autograd import _wgrad_ops_cublas
    return _wgrad_ops_cublas(ds, xs, tp)


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    cases = [([700, 0, 130, 5000, 63, 64, 1], 256, 128), ([4000, 2500, 1200], 1536, 512), ([3000, 3001], 512, 1024),
             ([1000, 900, 300, 80, 50, 20], 96, 72), ([72000, 48000, 24000, 8000, 5000, 3000], 1536, 512),
             ([72000, 48000, 24000, 8000, 5000, 3000], 512, 512)]
    for variant in [0]:
        for sizes, M, Nn in cases:
            tp = [0]
            for s in sizes:
                tp.append(tp[-1] + s)
            N = tp[-1]
            dy = torch.randn(N, M, device=dev) * torch.rand(N, 1, device=dev)
            x = torch.randn(N, Nn, device=dev)
            ds, xs = ops.to_operand(dy, ops.OPF_BF16X3), ops.to_operand(x, ops.OPF_BF16X3)
            got = ops.typed_wgrad(ds, xs, tp)
            torch.cuda.synchronize()
            want = ref(dy, x, tp)
            err = ((got.double() - want).abs().amax() / want.abs().amax()).item()
            rec = {"variant": variant, "sizes": sizes, "M": M, "Nn": Nn, "rel_err": err}
            if N > 50000 and variant == 0:
                for name, fn in (("tc_us", lambda: ops.typed_wgrad(ds, xs, tp)), ("cublas_us", lambda: cublas_route(ds, xs, tp))):
                    for _ in range(3):
                        fn()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for _ in range(10):
                        fn()
                    b.record()
                    torch.cuda.synchronize()
                    rec[name] = a.elapsed_time(b) * 100
                rec["tflops_alg"] = 2.0 * N * M * Nn / (rec["tc_us"] * 1e-6) / 1e12
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
