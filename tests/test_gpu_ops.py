"""GPU parity of the individual C-ABI kernels against the CPU oracle primitives (oracle/primitives.py)
and fp64 torch restatements, on seeded random inputs incl. the edge cases of the domain
(empty types / segments, ragged rows, hubs, zero in-degree)."""
import math

import pytest
import torch

from oracle import primitives as P
from wsi_hgnn_b200 import ops

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


# relative L2 tolerance of ONE tensor-core GEMM per operand format (fp64 reference): the 3-term bf16 split is fp32-grade;
# one fp16 / bf16 pass carries the operand rounding 2^-11 / 2^-8 (x sqrt(2)/sqrt(3) in the L2 norm of a random product)
PREC_TOL = {"bf16x3": 2e-5, "fp16": 6e-4, "bf16": 5e-3}

TC_SHAPES = [([600, 0, 700], 512, 1536), ([128, 256, 300], 200, 200), ([1000], 1024, 512), ([513], 64, 64),
             ([130, 1, 127, 500, 0, 3], 256, 520), ([2731, 2731, 2730], 512, 512)]


@pytest.fixture
def precision(request):
    prev = ops.set_matmul_precision(request.param)
    yield request.param
    ops.set_matmul_precision(prev)


@pytest.mark.parametrize("precision", ["bf16x3", "fp16", "bf16"], indirect=True)
@pytest.mark.parametrize("counts,K,n_out,impl",
                         [(c, k, n, i) for c, k, n in [([70, 0, 133], 48, 96), ([1, 2, 3], 7, 5), ([300], 200, 200),
                                                       ([128, 256], 512, 1536), ([0, 0, 5], 16, 256)]
                          for i in (ops.IMPL_SIMT, ops.IMPL_AUTO)] +
                         [(c, k, n, ops.IMPL_TC) for c, k, n in TC_SHAPES])
def test_typed_linear_epilogues(counts, K, n_out, impl, precision):
    if impl == ops.IMPL_SIMT and precision != "bf16x3":
        pytest.skip("the SIMT path is fp32 whatever the operand format")
    TOL = 2e-5 if impl == ops.IMPL_SIMT else PREC_TOL[precision]
    g = torch.Generator().manual_seed(sum(counts) + K)
    T = len(counts)
    ptr = [0]
    for c in counts:
        ptr.append(ptr[-1] + c)
    N = ptr[-1]
    x = torch.randn(N, K, generator=g)
    w = torch.randn(T, n_out, K, generator=g) / math.sqrt(K)
    b = torch.randn(T, n_out, generator=g)
    skip = torch.randn(T, generator=g)
    res = torch.randn(N, n_out, generator=g)
    mask = (torch.rand(N, n_out, generator=g) > 0.3).float() / 0.7
    gate = (torch.rand(N, generator=g) > 0.25).float()
    scale = (torch.rand(N, generator=g) > 0.5).float()

    def ref(act=False, use_skip=False, use_mask=False, use_gate=False, use_scale=False):
        out = torch.empty(N, n_out, dtype=torch.float64)
        for t in range(T):
            a, z = ptr[t], ptr[t + 1]
            v = x[a:z].double() @ w[t].double().T + b[t].double()
            if act:
                v = torch.nn.functional.gelu(v)
            if use_mask:
                v = v * mask[a:z].double()
            if use_skip:
                al = torch.sigmoid(skip[t].double())
                mixed = v * al + res[a:z].double() * (1 - al)
                v = torch.where(gate[a:z, None] != 0, mixed, res[a:z].double()) if use_gate else mixed
            if use_scale:
                v = v * scale[a:z, None].double()
            out[a:z] = v
        return out

    c = lambda t: t.cuda()
    y = ops.typed_linear(c(x), c(w), c(b), ptr, impl=impl)
    assert rel(y, ref()) < TOL
    y = ops.typed_linear(c(x), c(w), c(b), ptr, act=ops.ACT_GELU, impl=impl)
    assert rel(y, ref(act=True)) < TOL
    y = ops.typed_linear(c(x), c(w), c(b), ptr, skip=c(skip), res=c(res), drop_mask=c(mask), row_gate=c(gate),
                         row_scale=c(scale), impl=impl)
    assert rel(y, ref(use_skip=True, use_mask=True, use_gate=True, use_scale=True)) < TOL
    # strided input / output views (the K|V|Q fused buffer is sliced by column)
    big = torch.zeros(N, n_out + 8, device="cuda")
    ops.typed_linear(c(x), c(w), None, ptr, out=big[:, 8:], impl=impl)
    assert rel(big[:, 8:], ref() - torch.cat([b[t].double().expand(counts[t], n_out) for t in range(T)])) < TOL
    assert float(big[:, :8].abs().sum()) == 0.0


@pytest.mark.parametrize("precision", ["bf16x3", "fp16", "bf16"], indirect=True)
@pytest.mark.parametrize("counts,K,n_out", [([600, 0, 700], 512, 512), ([513], 64, 136), ([300, 300, 424], 1024, 256)])
def test_typed_linear_op_chain(counts, K, n_out, precision):
    """Operands already in operand form in, fp32 + operand-form result out (the operand of the next GEMM)."""
    TOL = PREC_TOL[precision]
    split = precision == "bf16x3"
    g = torch.Generator().manual_seed(K + n_out)
    T = len(counts)
    ptr = [0]
    for c in counts:
        ptr.append(ptr[-1] + c)
    N = ptr[-1]
    x = torch.randn(N, K, generator=g)
    w = torch.randn(T, n_out, K, generator=g) / math.sqrt(K)
    b = torch.randn(T, n_out, generator=g)
    skip = torch.randn(T, generator=g)
    res = torch.randn(N, n_out, generator=g)
    gate = (torch.rand(N, generator=g) > 0.25).float()
    xs = ops.to_operand(x.cuda())
    assert xs.shape == ((2 * N if split else N), K) and xs.dtype == (torch.float16 if precision == "fp16" else torch.bfloat16)
    back = (xs[:N].float() + xs[N:].float()) if split else xs.float()
    assert float((back.cpu() - x).abs().max() / x.abs().max()) < {"bf16x3": 2 ** -15, "fp16": 2 ** -10, "bf16": 2 ** -7}[precision]
    if precision == "fp16":                                  # values beyond the fp16 range are clamped, never inf
        big = ops.to_operand(torch.tensor([[1e6, -1e6, 3.0, 1e-9] * 2], device="cuda"))
        assert torch.isfinite(big.float()).all() and float(big[0, 0]) == 65504.0 and float(big[0, 1]) == -65504.0
    ws = ops.to_operand(w.cuda())
    ref = torch.empty(N, n_out, dtype=torch.float64)
    for t in range(T):
        a, z = ptr[t], ptr[t + 1]
        v = x[a:z].double() @ w[t].double().T + b[t].double()
        al = torch.sigmoid(skip[t].double())
        ref[a:z] = torch.where(gate[a:z, None] != 0, v * al + res[a:z].double() * (1 - al), res[a:z].double())
    y, ys = ops.typed_linear_op(xs, ws, b.cuda(), ptr, n_out, skip=skip.cuda(), res=res.cuda(), row_gate=gate.cuda(),
                                   want_op=True)
    assert rel(y, ref) < TOL
    assert rel((ys[:N].float() + ys[N:].float()) if split else ys.float(), ref) < max(TOL, {"bf16x3": 0, "fp16": 8e-4, "bf16": 6e-3}[precision])
    y2, none = ops.typed_linear_op(xs, ws, b.cuda(), ptr, n_out, skip=skip.cuda(), res=res.cuda(), row_gate=gate.cuda())
    assert none is None and torch.equal(y2, y)
    only, ys2 = ops.typed_linear_op(xs, ws, b.cuda(), ptr, n_out, skip=skip.cuda(), res=res.cuda(),
                                       row_gate=gate.cuda(), want_y=False, want_op=True)
    assert only is None and torch.equal(ys2, ys)


@pytest.mark.parametrize("counts,M,Nn", [([700, 0, 130, 5000, 63, 64, 1], 256, 128), ([4000, 2500, 1200], 1536, 512),
                                         ([3000, 3001], 512, 1024), ([1000, 900, 300, 80, 50, 20], 96, 72),
                                         ([20, 30, 600], 64, 64), ([30000, 9000, 1], 512, 512)])
def test_typed_wgrad(counts, M, Nn):
    """dW[t] = dY_t^T X_t on tcgen05 (MN-major operands, K-chunked over the rows of a type, SIMT tail) against fp64:
    ragged and empty types, types below one 64-row block, M / Nn below and across the 256-wide tile."""
    torch.manual_seed(len(counts) * 1000 + M + Nn)
    dev = torch.device("cuda", 0)
    tp = [0]
    for c in counts:
        tp.append(tp[-1] + c)
    N = tp[-1]
    assert ops.typed_wgrad_ok(N, M, Nn, len(counts))
    dy = torch.randn(N, M, device=dev) * torch.rand(N, 1, device=dev)
    x = torch.randn(N, Nn, device=dev)
    got = ops.typed_wgrad(ops.to_operand(dy, ops.OPF_BF16X3), ops.to_operand(x, ops.OPF_BF16X3), tp)
    want = torch.stack([dy[tp[t]:tp[t + 1]].double().t() @ x[tp[t]:tp[t + 1]].double() for t in range(len(counts))])
    assert got.shape == want.shape
    for t in range(len(counts)):
        if counts[t] == 0:
            assert not got[t].any()
        else:
            assert rel(got[t], want[t]) < 3e-5, (t, counts[t])


def test_typed_wgrad_refuses_unfit_shapes():
    assert not ops.typed_wgrad_ok(100, 512, 512, 3)          # too few rows
    assert not ops.typed_wgrad_ok(5000, 200, 512, 3)         # n_out not a multiple of 32
    assert not ops.typed_wgrad_ok(5000, 512, 100, 3)         # K not a multiple of 8
    dev = torch.device("cuda", 0)
    with pytest.raises(Exception):
        ops.typed_wgrad(torch.zeros(2000, 200, dtype=torch.bfloat16, device=dev),
                        torch.zeros(2000, 512, dtype=torch.bfloat16, device=dev), [0, 1000])


def test_typed_linear_tc_refuses_unfit_shapes():
    """impl=2 (force tcgen05) must fail loudly - never silently take another path."""
    x = torch.randn(8, 7, device="cuda")
    w = torch.randn(1, 5, 7, device="cuda")
    with pytest.raises(NotImplementedError):
        ops.typed_linear(x, w, None, [0, 8], impl=ops.IMPL_TC)


def _random_csr(n_dst, n_src, n_edges, n_rel, g, hub=0, isolated=0.2):
    dst = torch.randint(0, n_dst, (n_edges,), generator=g)
    iso = torch.rand(n_dst, generator=g) < isolated
    dst = dst[~iso[dst]]
    if hub:
        dst = torch.cat([dst, torch.zeros(hub, dtype=torch.int64)])
    E = dst.numel()
    src = torch.randint(0, n_src, (E,), generator=g)
    rel = torch.randint(0, n_rel, (E,), generator=g)
    sim = torch.rand(E, generator=g) * 2 - 1
    order = torch.argsort(dst * 256 + rel, stable=True)
    dst, src, rel, sim = dst[order], src[order], rel[order], sim[order]
    rowptr = torch.zeros(n_dst + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.bincount(dst, minlength=n_dst), 0)
    return rowptr, src, dst, rel, sim


def _heat_attn_ref(k, v, q, src, dst, rel, sim, inv_r, ew, eb, H, n_rel):
    """per relation: v_dot_u -> score -> edge_softmax -> u_mul_e/sum; then sum over relations * inv_r."""
    N, D = q.shape
    dk = D // H
    k3, v3, q3 = (t.double().view(-1, H, dk) for t in (k, v, q))
    out = torch.zeros(N, H, dk, dtype=torch.float64)
    attn = torch.zeros(src.numel(), H, dtype=torch.float64)
    for r in range(n_rel):
        m = torch.nonzero(rel == r).reshape(-1)
        if m.numel() == 0:
            continue
        t = P.v_dot_u(q3, k3, src[m], dst[m])
        score = t.sum(-1) * (ew * sim[m].double() + eb).view(-1, 1) / math.sqrt(dk)
        a = P.edge_softmax(score, dst[m], N)
        attn[m] = a
        out += P.u_mul_e_sum(v3, a.unsqueeze(-1), src[m], dst[m], N)
    return (out * inv_r.double().view(-1, 1, 1)).view(N, D), attn


@pytest.mark.parametrize("D,H,perm", [(128, 4, True), (256, 8, True), (512, 4, True), (512, 1, True),
                                      (1024, 32, True), (384, 2, True), (128, 4, False), (200, 4, False),
                                      (64, 4, False), (96, 3, False), (512, 4, False), (32, 4, False)])
def test_hetero_attn_fwd(D, H, perm):
    g = torch.Generator().manual_seed(D * 7 + H)
    n_dst, n_rel = 257, 5
    rowptr, src, dst, rel, sim = _random_csr(n_dst, n_dst, 1500, n_rel, g, hub=150)
    k = torch.randn(n_dst, D, generator=g)
    v = torch.randn(n_dst, D, generator=g)
    q = torch.randn(n_dst, D, generator=g) * 0.5
    inv_r = torch.full((n_dst,), 1.0 / n_rel)
    inv_r[torch.rand(n_dst, generator=g) < 0.1] = 0.0          # passthrough rows
    ew, eb = 0.9, -0.3
    ref, attn_ref = _heat_attn_ref(k, v, q, src, dst, rel, sim, inv_r, ew, eb, H, n_rel)
    kvq = torch.cat([k, v, q], 1)
    if perm:
        p = ops.head_perm(D, H)
        kvq = torch.cat([k[:, p], v[:, p], q[:, p]], 1)
    kvq = kvq.cuda()
    c = lambda t, dt: t.to(dt).cuda()
    agg, attn = ops.hetero_attn(kvq[:, :D], kvq[:, D:2 * D], kvq[:, 2 * D:], c(rowptr, torch.int32),
                                c(src, torch.int32), c(sim, torch.float32), c(rel, torch.uint8), inv_r.cuda(),
                                torch.tensor([[ew]]).cuda(), torch.tensor([eb]).cuda(), D, H, perm, want_attn=True)
    # without the attention output the lane-grouped layout takes the TMA-ring kernel (whole rows, hub of 150+ edges)
    agg2 = ops.hetero_attn(kvq[:, :D], kvq[:, D:2 * D], kvq[:, 2 * D:], c(rowptr, torch.int32),
                           c(src, torch.int32), c(sim, torch.float32), c(rel, torch.uint8), inv_r.cuda(),
                           torch.tensor([[ew]]).cuda(), torch.tensor([eb]).cuda(), D, H, perm)
    assert rel_ok(agg2, agg, 1e-5)
    agg = agg.cpu()
    if perm:
        un = torch.empty_like(agg)
        un[:, p] = agg
        agg = un
    assert rel_ok(agg, ref, 2e-5)
    live = inv_r[dst] != 0
    assert rel_ok(attn.cpu()[live], attn_ref[live], 2e-5)
    assert float(agg[inv_r == 0].abs().sum()) == 0.0


@pytest.mark.parametrize("D,H,chunk", [(512, 4, 16), (128, 4, 4), (256, 8, 1), (1024, 32, 16), (384, 2, 7), (512, 1, 64)])
def test_hetero_attn_work_list(D, H, chunk):
    """Hub-balanced work list (chunks of split rows + merge launch) == whole-row reference."""
    from wsi_hgnn_b200.hetero_graph import GraphPlan
    g = torch.Generator().manual_seed(D + H + chunk)
    n_dst, n_rel = 301, 5
    rowptr, src, dst, rel, sim = _random_csr(n_dst, n_dst, 1800, n_rel, g, hub=200)
    k = torch.randn(n_dst, D, generator=g)
    v = torch.randn(n_dst, D, generator=g)
    q = torch.randn(n_dst, D, generator=g) * 0.5
    inv_r = torch.full((n_dst,), 1.0 / n_rel)
    inv_r[torch.rand(n_dst, generator=g) < 0.1] = 0.0
    inv_r[0] = 1.0 / n_rel                                       # the hub row stays live
    ew, eb = -0.7, 0.2
    ref, _ = _heat_attn_ref(k, v, q, src, dst, rel, sim, inv_r, ew, eb, H, n_rel)
    p = ops.head_perm(D, H)
    kvq = torch.cat([k[:, p], v[:, p], q[:, p]], 1).cuda()
    plan = GraphPlan()
    plan.N, plan.E, plan.device = n_dst, int(src.numel()), torch.device("cuda")
    plan.rowptr = rowptr.to(torch.int32).cuda()
    plan.e_rel = rel.to(torch.uint8).cuda()
    work = plan.attn_work(chunk)
    assert work["n_split"] > 0 and work["n_part"] > work["n_split"]
    agg = ops.hetero_attn_work(kvq[:, :D], kvq[:, D:2 * D], kvq[:, 2 * D:], work, src.to(torch.int32).cuda(),
                               sim.float().cuda(), plan.e_rel, inv_r.cuda(), torch.tensor([[ew]]).cuda(),
                               torch.tensor([eb]).cuda(), D, H).cpu()
    un = torch.empty_like(agg)
    un[:, p] = agg
    assert rel_ok(un, ref, 2e-5)
    assert float(un[inv_r == 0].abs().sum()) == 0.0
    # operand-form outputs of the same result: bf16 [hi; lo], fp16, bf16
    for opf, dt, rows, eps in ((ops.OPF_BF16X3, torch.bfloat16, 2 * n_dst, 2 ** -15), (ops.OPF_F16, torch.float16, n_dst, 2 ** -10),
                               (ops.OPF_BF16, torch.bfloat16, n_dst, 2 ** -7)):
        sp = ops.hetero_attn_work(kvq[:, :D], kvq[:, D:2 * D], kvq[:, 2 * D:], work, src.to(torch.int32).cuda(),
                                  sim.float().cuda(), plan.e_rel, inv_r.cuda(), torch.tensor([[ew]]).cuda(),
                                  torch.tensor([eb]).cuda(), D, H, op_out=True, opf=opf)
        assert sp.dtype == dt and sp.shape == (rows, D)
        back = ((sp[:n_dst].float() + sp[n_dst:].float()) if opf == ops.OPF_BF16X3 else sp.float()).cpu()
        assert float((back - agg).abs().max()) <= eps * float(agg.abs().max())


@pytest.mark.parametrize("kernel", [2])
@pytest.mark.parametrize("D,H,chunk,kv_dt,q_dt", [(512, 4, 16, torch.float32, torch.float32), (512, 4, 16, torch.bfloat16, torch.bfloat16),
                                                  (256, 8, 4, torch.float16, torch.float32), (128, 4, 1, torch.float16, torch.float16),
                                                  (512, 1, 64, torch.bfloat16, torch.float32)])
def test_hetero_attn_ring_kernel(kernel, D, H, chunk, kv_dt, q_dt):
    """The TMA-ring kernel (the one large K | V footprints get), forced through the development knob, == the
    register-path kernel on the same (rounded) inputs: hub chunks with the fused merge, the self-re-arming work queue,
    zero in-degree and passthrough rows, 16-bit K / V and q storage."""
    from wsi_hgnn_b200.hetero_graph import GraphPlan
    g = torch.Generator().manual_seed(D + H + chunk)
    n_dst, n_rel = 301, 5
    rowptr, src, dst, rel_, sim = _random_csr(n_dst, n_dst, 1800, n_rel, g, hub=200)
    p = ops.head_perm(D, H)
    k = torch.randn(n_dst, D, generator=g)[:, p].cuda().to(kv_dt)
    v = torch.randn(n_dst, D, generator=g)[:, p].cuda().to(kv_dt)
    kv = torch.cat([k, v], 1).contiguous()
    q = (torch.randn(n_dst, D, generator=g) * 0.5)[:, p].cuda().to(q_dt).contiguous()
    inv_r = torch.full((n_dst,), 1.0 / n_rel)
    inv_r[torch.rand(n_dst, generator=g) < 0.1] = 0.0
    inv_r[0] = 1.0 / n_rel
    plan = GraphPlan()
    plan.N, plan.E, plan.device = n_dst, int(src.numel()), torch.device("cuda")
    plan.rowptr = rowptr.to(torch.int32).cuda()
    plan.e_rel = rel_.to(torch.uint8).cuda()
    work = plan.attn_work(chunk)
    args = (kv[:, :D], kv[:, D:], q, work, src.to(torch.int32).cuda(), sim.float().cuda(), plan.e_rel, inv_r.cuda(),
            torch.tensor([[-0.7]]).cuda(), torch.tensor([0.2]).cuda(), D, H)
    try:
        ops.dev_set("attn_kernel", 1)
        want = ops.hetero_attn_work(*args)
        ops.dev_set("attn_kernel", kernel)
        got = ops.hetero_attn_work(*args)
        got2 = ops.hetero_attn_work(*args)                       # the queue / arrival counters re-arm themselves
        got_op = ops.hetero_attn_work(*args, op_out=True, opf=ops.OPF_F16)
    finally:
        ops.dev_set("attn_kernel", 0)
    assert rel(got, want) < 2e-6 and torch.equal(got, got2)
    assert float(got[inv_r.cuda() == 0].abs().sum()) == 0.0
    assert rel(got_op.float(), want) < 1e-3


def rel_ok(a, b, tol):
    e = rel(a, b)
    assert e < tol, f"rel err {e:.3e}"
    return True


@pytest.mark.parametrize("op", ["sum", "mean", "max"])
@pytest.mark.parametrize("D", [512, 200, 5])
def test_segment_pool(op, D):
    g = torch.Generator().manual_seed(D)
    lens = [0, 17, 1, 0, 300, 2500, 0, 64]
    ptr = [0]
    for n in lens:
        ptr.append(ptr[-1] + n)
    x = torch.randn(ptr[-1], D, generator=g)
    ref = P.segment_readout(x.double(), torch.tensor(lens), op)
    out = ops.segment_pool(x.cuda(), torch.tensor(ptr, dtype=torch.int32).cuda(), len(lens), op)
    assert rel(out, ref) < 1e-5
    # one huge segment: goes through the split + finish path
    x = torch.randn(40000, D, generator=g)
    ref = P.segment_readout(x.double(), torch.tensor([40000]), op)
    out = ops.segment_pool(x.cuda(), torch.tensor([0, 40000], dtype=torch.int32).cuda(), 1, op)
    assert rel(out, ref) < 1e-5


@pytest.mark.parametrize("op", ["sum", "mean", "max"])
@pytest.mark.parametrize("T,B,D,n_out", [(3, 1, 512, 2), (2, 5, 200, 8), (6, 3, 128, 1)])
def test_segment_pool_affine(op, T, B, D, n_out):
    """fused typed readout + narrow affine prediction == pool, per-type linear, sum over types."""
    g = torch.Generator().manual_seed(T * 100 + B * 10 + n_out)
    lens = torch.randint(0, 400, (T * B,), generator=g)
    lens[1 % (T * B)] = 0                                   # an empty (type, graph) segment
    lens[0] = 5000                                          # one long segment (several row slabs)
    ptr = torch.zeros(T * B + 1, dtype=torch.int64)
    ptr[1:] = torch.cumsum(lens, 0)
    x = torch.randn(int(ptr[-1]), D, generator=g)
    M = torch.randn(T, n_out, D, generator=g) / math.sqrt(D)
    c = torch.randn(T, n_out, generator=g)
    bt = torch.randn(n_out, generator=g)
    scale = (lens > 0).float()
    pooled = P.segment_readout(x.double(), lens, op)                          # [T*B, D]
    ref = bt.double() + sum(scale.view(T, B)[t].double()[:, None] *
                            (pooled.view(T, B, D)[t] @ M[t].double().T + c[t].double()) for t in range(T))
    out = ops.segment_pool_affine(x.cuda(), ptr.to(torch.int32).cuda(), T, B, op, M.cuda(), c.cuda(), bt.cuda(), scale.cuda())
    assert rel(out, ref) < 1e-5
    out2 = ops.segment_pool_affine(x.cuda(), ptr.to(torch.int32).cuda(), T, B, op, M.cuda(), c.cuda(), bt.cuda(), scale.cuda(),
                                   out=out.clone(), accumulate=True)
    assert rel(out2, 2 * ref) < 1e-5


def test_typed_layernorm():
    g = torch.Generator().manual_seed(1)
    counts, D = [33, 0, 80], 200
    ptr = [0, 33, 33, 113]
    x = torch.randn(113, D, generator=g) * 3 + 1
    gamma = torch.randn(3, D, generator=g)
    beta = torch.randn(3, D, generator=g)
    ref = torch.cat([torch.nn.functional.layer_norm(x[ptr[t]:ptr[t + 1]].double(), (D,), gamma[t].double(),
                                                    beta[t].double(), 1e-5) for t in range(3)])
    y = ops.typed_layernorm(x.cuda(), gamma.cuda(), beta.cuda(), ptr)
    assert rel(y, ref) < 1e-5


@pytest.mark.parametrize("D", [128, 512, 1024])
def test_typed_layernorm_vector_path_with_operand_output(D):
    """D % 128 == 0: the register-resident kernel; row gate (passthrough rows), in place, and the operand-form copy of
    the result == the conversion of the result."""
    g = torch.Generator().manual_seed(D)
    ptr = [0, 33, 33, 113]
    x = torch.randn(113, D, generator=g) * 3 + 1
    gamma, beta = torch.randn(3, D, generator=g), torch.randn(3, D, generator=g)
    gate = (torch.rand(113, generator=g) > 0.3).float()
    ln = torch.cat([torch.nn.functional.layer_norm(x[ptr[t]:ptr[t + 1]].double(), (D,), gamma[t].double(),
                                                   beta[t].double(), 1e-5) for t in range(3)])
    ref = torch.where(gate.unsqueeze(1) != 0, ln, x.double())
    for opf in (ops.OPF_BF16X3, ops.OPF_F16, ops.OPF_BF16):
        xc = x.cuda()
        y, y_op = ops.typed_layernorm(xc, gamma.cuda(), beta.cuda(), ptr, inplace=True, row_gate=gate.cuda(), op_out=True, opf=opf)
        assert y.data_ptr() == xc.data_ptr() and rel(y, ref) < 1e-5
        want = ops.to_operand(y, opf)
        assert torch.equal(y_op.view(torch.int16), want.view(torch.int16))


def test_to_operand_colsum():
    """Operand conversion fused with the per-type column sums (bias gradient) == the two separate kernels."""
    g = torch.Generator().manual_seed(9)
    counts = [700, 0, 255, 256, 257, 3]
    tp = [0]
    for c in counts:
        tp.append(tp[-1] + c)
    big = torch.randn(tp[-1], 520, generator=g).cuda()
    x = big[:, :512]                                       # row-strided view
    xs, cs = ops.to_operand_colsum(x, tp)
    assert torch.equal(xs.view(torch.int16), ops.to_operand(x.contiguous(), ops.OPF_BF16X3).view(torch.int16))
    ref = torch.stack([x[tp[t]:tp[t + 1]].double().sum(0) for t in range(len(counts))])
    assert rel(cs, ref) < 1e-6 and not cs[1].any()
    assert torch.equal(cs, ops.to_operand_colsum(x, tp)[1])            # deterministic


def test_gather_rows16():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(300, 256, generator=g).cuda().to(torch.bfloat16)
    idx = torch.randint(0, 300, (1000,), generator=g).to(torch.int32).cuda()
    assert torch.equal(ops.gather_rows16(x, idx), x[idx.long()])


@pytest.mark.parametrize("dk,H", [(50, 4), (128, 4), (32, 8)])
def test_rel_transform(dk, H):
    g = torch.Generator().manual_seed(dk)
    R, S, Nrows = 6, 300, 120
    seg_rel = torch.randint(0, R, (S,), generator=g)
    seg_rel[seg_rel == 2] = 3                       # an empty relation group
    seg_row = torch.randint(0, Nrows, (S,), generator=g)
    x = torch.randn(Nrows, H * dk, generator=g)
    w = torch.randn(R, H, dk, dk, generator=g) / math.sqrt(dk)
    order = torch.argsort(seg_rel, stable=True)
    rel_ptr = [0] + torch.cumsum(torch.bincount(seg_rel, minlength=R), 0).tolist()
    xs = x[seg_row].double().view(S, H, dk)
    for w_kn in (False, True):
        eq = "shk,shkn->shn" if w_kn else "shk,shnk->shn"
        ref = torch.einsum(eq, xs, w[seg_rel].double()).reshape(S, H * dk)
        y = ops.rel_transform(x.cuda(), seg_row[order].to(torch.int32).cuda(), order.to(torch.int32).cuda(), w.cuda(),
                              ops.host_i32(rel_ptr), R, H, dk, w_kn, S)
        assert rel(y, ref) < 2e-5


@pytest.mark.parametrize("precision", ["bf16x3", "fp16", "bf16"], indirect=True)
def test_gather_to_operand(precision):
    """Row gather fused into the fp32 -> operand-form conversion == conversion of the gathered matrix."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(500, 256, generator=g).cuda()
    idx = torch.randint(0, 500, (1300,), generator=g).to(torch.int32).cuda()
    got = ops.gather_to_operand(x, idx)
    want = ops.to_operand(x[idx.long()].contiguous())
    assert got.dtype == want.dtype and torch.equal(got.view(torch.int16), want.view(torch.int16))


@pytest.mark.parametrize("kv_dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_hetero_attn_seg_work_list(kv_dtype):
    """HGT segment attention through the work list (segments visited in another order than the edge order, lane-grouped
    layout, operand-form output) == the plain segment kernel."""
    g = torch.Generator().manual_seed(11)
    D, H, N, S = 256, 8, 400, 900
    lens = torch.randint(1, 6, (S,), generator=g)
    lens[5] = 70                                   # one long segment
    seg_ptr = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(lens, 0)])
    E = int(seg_ptr[-1])
    e_src = torch.randint(0, N, (E,), generator=g).to(torch.int32).cuda()
    R = 5
    seg_rel = torch.randint(0, R, (S,), generator=g).to(torch.int32)
    pri = (torch.rand(R, H, generator=g) + 0.5).cuda()
    k, v = torch.randn(N, D, generator=g).cuda(), torch.randn(N, D, generator=g).cuda()
    qseg = torch.randn(S, D, generator=g).cuda()
    ref = ops.hetero_attn_seg(k, v, qseg, seg_ptr.to(torch.int32).cuda(), seg_rel.cuda(), e_src, pri, D, H)   # natural order
    pm = ops.head_perm(D, H).cuda()
    order = torch.randperm(S, generator=g)
    items = torch.stack([torch.arange(S), seg_ptr[:-1][order], seg_ptr[1:][order], torch.full((S,), -1)], 1).to(torch.int32).cuda()
    kp, vp = k[:, pm].to(kv_dtype).contiguous(), v[:, pm].to(kv_dtype).contiguous()
    out, out_op = ops.hetero_attn_seg(kp, vp, qseg[order.cuda()][:, pm].contiguous(), None, seg_rel[order].contiguous().cuda(), e_src,
                                      pri, D, H, True, items=items, want_out=True, op_out=True, opf=ops.OPF_BF16X3)
    inv = torch.empty_like(pm)
    inv[pm] = torch.arange(D, device="cuda")
    got = torch.empty_like(out)
    got[order.cuda()] = out[:, inv]
    tol = 2e-5 if kv_dtype == torch.float32 else (3e-3 if kv_dtype == torch.float16 else 2e-2)
    assert rel(got, ref) < tol
    assert rel(out_op[:S].float() + out_op[S:].float(), out) < 1e-5


def test_segment_combine_indexed():
    g = torch.Generator().manual_seed(5)
    N, D = 300, 256
    cnt = torch.randint(0, 4, (N,), generator=g)
    rsp = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(cnt, 0)])
    S = int(rsp[-1])
    msg = torch.randn(S, D, generator=g)
    inv_r = torch.where(cnt > 0, 1.0 / cnt.clamp_min(1).float(), torch.zeros(N))
    pos = torch.randperm(S, generator=g)
    shuffled = torch.empty_like(msg)
    shuffled[pos] = msg
    ref = torch.stack([msg[rsp[i]:rsp[i + 1]].double().sum(0) * float(inv_r[i]) for i in range(N)])
    agg, agg_op = ops.segment_combine(shuffled.cuda(), rsp.to(torch.int32).cuda(), inv_r.cuda(), N, D,
                                      seg_pos=pos.to(torch.int32).cuda(), op_out=True, opf=ops.OPF_F16)
    assert rel(agg, ref) < 1e-6
    assert rel(agg_op.float(), ref) < 1e-3
    plain = ops.segment_combine(msg.cuda(), rsp.to(torch.int32).cuda(), inv_r.cuda(), N, D)
    assert torch.equal(plain, agg)


def test_ops_reject_cpu_tensors():
    with pytest.raises(RuntimeError):
        ops.typed_linear(torch.zeros(4, 4), torch.zeros(1, 4, 4), None, [0, 4])
