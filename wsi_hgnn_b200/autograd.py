"""torch.autograd wiring of the C-ABI kernels: what makes `loss.backward()` of the reference's training step
(trainer/train_gnn.py:55-79) work on the CUDA path.  Forward and backward both run libwsi_hgnn.so kernels, including the
weight gradient dW_t = dY_t^T X_t (wsi_typed_wgrad, tcgen05 on MN-major operands); shapes that kernel does not take
(n_out not a multiple of 32, fewer than 512 rows: the readout heads) fall to a cuBLAS GEMM through torch.mm."""
from typing import Sequence

import torch

from . import ops


_MM_OUT_DTYPE = [None]      # does torch.mm(bf16, bf16, out_dtype=fp32) exist in this build?


def _wgrad(dy: torch.Tensor, x: torch.Tensor, tp: Sequence[int]) -> torch.Tensor:
    """dW[t] = dY_t^T X_t, a plain dense GEMM per node type -> library call (cuBLAS).  fp32-accurate at tensor-core
    speed: both operands go through the product's own fp32 -> bf16 [hi; lo] split kernel and the product is
    hi.hi + hi.lo + lo.hi accumulated in fp32 (the same 3-term scheme as the forward GEMM); plain fp32 cuBLAS when this
    torch build has no mm(out_dtype=)."""
    T = len(tp) - 1
    N = int(tp[-1])
    if _MM_OUT_DTYPE[0] is not False and N > 0 and x.shape[1] % 8 == 0 and dy.shape[1] % 8 == 0:
        try:
            xs, ds = ops.to_operand(x, ops.OPF_BF16X3), ops.to_operand(dy, ops.OPF_BF16X3)
            out = []
            for t in range(T):
                a, z = tp[t], tp[t + 1]
                dh, dl, xh, xl = ds[a:z].t(), ds[N + a:N + z].t(), xs[a:z], xs[N + a:N + z]
                w = torch.mm(dh, xh, out_dtype=torch.float32)
                w += torch.mm(dh, xl, out_dtype=torch.float32)
                w += torch.mm(dl, xh, out_dtype=torch.float32)
                out.append(w)
            _MM_OUT_DTYPE[0] = True
            return torch.stack(out)
        except (TypeError, NotImplementedError, RuntimeError):
            if _MM_OUT_DTYPE[0] is True:
                raise
            _MM_OUT_DTYPE[0] = False
    return torch.stack([dy[tp[t]:tp[t + 1]].t() @ x[tp[t]:tp[t + 1]] for t in range(T)])


def _wgrad_ops(ds: torch.Tensor, xs: torch.Tensor, tp: Sequence[int], type_ptr_c=None) -> torch.Tensor:
    """dW[t] = dY_t^T X_t from operands ALREADY in the [hi; lo] split form (no re-conversion): hi.hi + hi.lo + lo.hi,
    fp32 accumulate, on tcgen05 (wsi_typed_wgrad: the operands are read MN-major, nothing is transposed in memory)."""
    if ops.typed_wgrad_ok(int(tp[-1]), int(ds.shape[1]), int(xs.shape[1]), len(tp) - 1):
        return ops.typed_wgrad(ds, xs, tp, type_ptr_c)
    return _wgrad_ops_cublas(ds, xs, tp)


def _wgrad_ops_cublas(ds: torch.Tensor, xs: torch.Tensor, tp: Sequence[int]) -> torch.Tensor:
    """The same product for shapes wsi_typed_wgrad does not take (n_out not a multiple of 32): per-type cuBLAS bf16
    GEMMs with fp32 output on strided views of the operand planes."""
    T = len(tp) - 1
    N = int(tp[-1])
    if _MM_OUT_DTYPE[0] is not False:
        try:
            out = []
            for t in range(T):
                a, z = tp[t], tp[t + 1]
                dh, dl, xh, xl = ds[a:z].t(), ds[N + a:N + z].t(), xs[a:z], xs[N + a:N + z]
                w = torch.mm(dh, xh, out_dtype=torch.float32)
                w += torch.mm(dh, xl, out_dtype=torch.float32)
                w += torch.mm(dl, xh, out_dtype=torch.float32)
                out.append(w)
            _MM_OUT_DTYPE[0] = True
            return torch.stack(out)
        except (TypeError, NotImplementedError, RuntimeError):
            if _MM_OUT_DTYPE[0] is True:
                raise
            _MM_OUT_DTYPE[0] = False                     # this torch has no mm(out_dtype=): fp32 cuBLAS on the recombined planes
    df, xf = ds[:N].float() + ds[N:].float(), xs[:N].float() + xs[N:].float()
    return torch.stack([df[tp[t]:tp[t + 1]].t() @ xf[tp[t]:tp[t + 1]] for t in range(T)])


def _tc_chain(N: int, K: int, n_out: int) -> bool:
    return N > 0 and ops.train_opf() == ops.OPF_BF16X3 and ops.tc_ok(N, K, n_out) and ops.tc_ok(N, n_out, K)


class TypedLinearFn(torch.autograd.Function):
    """y[rows of type t] = x[rows of type t] @ w[t].T + b[t]   (reference: the per-node-type nn.Linear calls).
    On tensor-core shapes every operand is converted to the [hi; lo] split form ONCE: x for the forward GEMM and the
    weight gradient, dy for the data gradient and the weight gradient."""

    @staticmethod
    def forward(ctx, x, w, b, type_ptr: Sequence[int], type_ptr_c):
        x = x.contiguous()
        w = w.contiguous()
        tp = list(type_ptr)
        N, K, n_out = int(tp[-1]), int(w.shape[2]), int(w.shape[1])
        ctx.type_ptr, ctx.type_ptr_c, ctx.has_bias = tp, type_ptr_c, b is not None
        ctx.chain = _tc_chain(N, K, n_out)
        if ctx.chain:
            xs = ops.to_operand(x, ops.OPF_BF16X3)
            ctx.save_for_backward(xs, w)
            y, _ = ops.typed_linear_op(xs, ops.to_operand(w, ops.OPF_BF16X3), b.contiguous() if b is not None else None, tp,
                                       n_out, type_ptr_c=type_ptr_c, opf=ops.OPF_BF16X3)
            return y
        ctx.save_for_backward(x, w)
        return ops.typed_linear(x, w, b.contiguous() if b is not None else None, type_ptr, type_ptr_c=type_ptr_c,
                                opf=ops.train_opf())

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        tp = ctx.type_ptr
        dx = dw = db = None
        if ctx.chain:
            if ctx.has_bias and ctx.needs_input_grad[2] and dy.shape[1] % 8 == 0:
                ds, db = ops.to_operand_colsum(dy, tp, ctx.type_ptr_c)      # operand form + bias gradient: one read of dy
            else:
                ds = ops.to_operand(dy, ops.OPF_BF16X3)
            if ctx.needs_input_grad[0]:                  # dgrad: the same typed GEMM with W^T
                wt = ops.to_operand(w.transpose(1, 2).contiguous(), ops.OPF_BF16X3)
                dx, _ = ops.typed_linear_op(ds, wt, None, tp, int(w.shape[2]), type_ptr_c=ctx.type_ptr_c, opf=ops.OPF_BF16X3)
            if ctx.needs_input_grad[1]:
                dw = _wgrad_ops(ds, x, tp, ctx.type_ptr_c)
        else:
            if ctx.needs_input_grad[0]:
                # gradients keep the 3-term split (fp32 range and ~2^-17 accuracy): a single fp16 pass would need loss scaling
                dx = ops.typed_linear(dy, w.transpose(1, 2).contiguous(), None, tp, type_ptr_c=ctx.type_ptr_c, opf=ops.OPF_BF16X3)
            if ctx.needs_input_grad[1]:                  # wgrad: plain dense GEMM per node type (cuBLAS)
                dw = _wgrad(dy, x, tp)
        if ctx.has_bias and ctx.needs_input_grad[2] and db is None:
            db = ops.typed_colsum(dy, tp) if dy.shape[1] % 4 == 0 and dy.shape[0] > 0 else \
                torch.stack([dy[tp[t]:tp[t + 1]].sum(0) for t in range(len(tp) - 1)])
        return dx, dw, db, None, None


class ALinearSkipFn(torch.autograd.Function):
    """out = drop(agg W_a^T + b_a) * sigma(skip_t) + x * (1 - sigma(skip_t)), passthrough rows keep x
    (models/HEATNet4.py:121-136) as ONE fused GEMM (the inference epilogue, with the dropout mask) in the forward and ONE
    row kernel (wsi_skip_mix_bwd) + dgrad / wgrad in the backward.  `lin` is never materialised: d out / d sigma(skip) =
    drop(lin) - x = (out - x) / sigma(skip)."""

    @staticmethod
    def forward(ctx, agg, w, b, x, skip_t, mask, gate, type_ptr: Sequence[int], type_ptr_c):
        tp = list(type_ptr)
        agg, w, x = agg.contiguous(), w.contiguous(), x.contiguous()
        ags = ops.to_operand(agg, ops.OPF_BF16X3)
        out, _ = ops.typed_linear_op(ags, ops.to_operand(w, ops.OPF_BF16X3), b.contiguous(), tp, int(w.shape[1]),
                                     skip=skip_t.contiguous(), res=x, drop_mask=mask, row_gate=gate, type_ptr_c=type_ptr_c,
                                     opf=ops.OPF_BF16X3)
        ctx.save_for_backward(ags, w, x, out, skip_t, mask if mask is not None else out.new_zeros(0), gate)
        ctx.type_ptr, ctx.type_ptr_c, ctx.has_mask = tp, type_ptr_c, mask is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        ags, w, x, out, skip_t, mask, gate = ctx.saved_tensors
        tp = ctx.type_ptr
        d_lin, d_x, d_alpha = ops.skip_mix_bwd(dout.contiguous(), out, x, mask if ctx.has_mask else None, skip_t.contiguous(),
                                               gate, tp, ctx.type_ptr_c)
        alpha = torch.sigmoid(skip_t)
        d_skip = d_alpha * alpha * (1 - alpha)
        ds, db = ops.to_operand_colsum(d_lin, tp, ctx.type_ptr_c)          # operand form + bias gradient: one read of d_lin
        wt = ops.to_operand(w.transpose(1, 2).contiguous(), ops.OPF_BF16X3)
        d_agg, _ = ops.typed_linear_op(ds, wt, None, tp, int(w.shape[2]), type_ptr_c=ctx.type_ptr_c, opf=ops.OPF_BF16X3)
        dw = _wgrad_ops(ds, ags, tp, ctx.type_ptr_c)
        return d_agg, dw, db, d_x, d_skip, None, None, None, None


class HeteroAttnFn(torch.autograd.Function):
    """agg = edge attention of one HEAT layer over all relations (models/HEATNet4.py:103-119); kvq [N, 3D] = K | V | Q in
    the lane-grouped column order."""

    @staticmethod
    def forward(ctx, kvq, e_w, e_b, plan, D: int, H: int):
        ctx.save_for_backward(kvq, e_w, e_b)
        ctx.plan, ctx.D, ctx.H = plan, D, H
        return ops.hetero_attn_work(kvq[:, :D], kvq[:, D:2 * D], kvq[:, 2 * D:], plan.attn_work(), plan.e_src, plan.e_sim,
                                    plan.e_rel, plan.node_inv_r, e_w, e_b, D, H)

    @staticmethod
    def backward(ctx, d_agg):
        kvq, e_w, e_b = ctx.saved_tensors
        plan, D, H = ctx.plan, ctx.D, ctx.H
        # two-pass backward on the transposed edge list (built once per plan): dK / dV rows are written once by their
        # source row's warp - no atomics, no zero fill of the [N, 3D] gradient
        if "transposed" not in plan.cache:
            plan.cache["transposed"] = ops.transposed_edges(plan.rowptr, plan.e_src, plan.N)
        d_kvq = torch.empty_like(kvq)
        d_e = ops.hetero_attn_bwd(kvq[:, :D], kvq[:, D:2 * D], kvq[:, 2 * D:], plan.rowptr, plan.e_src, plan.e_sim,
                                  plan.e_rel, plan.node_inv_r, e_w, e_b, D, H, d_agg.contiguous(), d_kvq[:, :D],
                                  d_kvq[:, D:2 * D], d_kvq[:, 2 * D:], row_order=plan.rows_by_degree(),
                                  transposed=plan.cache["transposed"])
        return d_kvq, d_e[0].reshape(e_w.shape), d_e[1].reshape(e_b.shape), None, None, None


class SegmentPoolFn(torch.autograd.Function):
    """Typed readout dgl.readout.{mean,sum}_nodes(graph, 'h', ntype=) (pooling/avg_pooling.py, sum_pooling.py)."""

    @staticmethod
    def forward(ctx, x, plan, n_seg: int, op: str):
        ctx.plan, ctx.op, ctx.n_seg = plan, op, n_seg
        pooled = ops.segment_pool(x, plan.seg_ptr, n_seg, op)
        if op == "max":
            ctx.save_for_backward(x, pooled)
        return pooled

    @staticmethod
    def backward(ctx, d_pooled):
        plan = ctx.plan
        seg_of_row, inv_n = plan.row_segments()
        d = d_pooled.index_select(0, seg_of_row)
        if ctx.op == "mean":
            d = d * inv_n.unsqueeze(1)
        elif ctx.op == "max":
            # dgl.readout.max_nodes / torch.max(dim=0): the gradient goes to ONE row per (segment, column), the first that
            # attains the maximum (ties have measure zero for real features; resolved like torch's max(0).indices)
            x, pooled = ctx.saved_tensors
            eq = x == pooled.index_select(0, seg_of_row)
            c = eq.to(torch.int64).cumsum(0)
            seg_first = plan.seg_ptr[:-1].to(torch.int64)                      # first row of every segment
            base = torch.zeros((ctx.n_seg, x.shape[1]), dtype=torch.int64, device=x.device)
            has_prev = seg_first > 0
            base[has_prev] = c.index_select(0, (seg_first[has_prev] - 1).clamp_max(max(x.shape[0] - 1, 0)))
            first = eq & ((c - base.index_select(0, seg_of_row)) == 1)
            d = d * first
        return d, None, None, None


class SkipMixFn(torch.autograd.Function):
    """out = lin * sigma(skip_t) + x * (1 - sigma(skip_t)) per node type; rows whose type has no incoming relation keep x
    (models/HEATNet4.py:122-136 incl. the KeyError passthrough :129-133).  Written as one Function because the
    generic autograd of a per-row gather of sigma(skip) is an atomic scatter of N values into T addresses."""

    @staticmethod
    def forward(ctx, lin, x, skip_t, type_ptr: Sequence[int], gate):
        T = len(type_ptr) - 1
        counts = torch.tensor([type_ptr[t + 1] - type_ptr[t] for t in range(T)], device=x.device)
        alpha_t = torch.sigmoid(skip_t)
        a_row = torch.repeat_interleave(alpha_t, counts).unsqueeze(1) * (gate != 0).unsqueeze(1)   # 0 => passthrough
        ctx.save_for_backward(lin, x, a_row, alpha_t)
        ctx.type_ptr = list(type_ptr)
        return lin * a_row + x * (1 - a_row)

    @staticmethod
    def backward(ctx, dout):
        lin, x, a_row, alpha_t = ctx.saved_tensors
        tp = ctx.type_ptr
        d_lin = dout * a_row
        d_x = dout * (1 - a_row)
        s_row = (dout * (lin - x)).sum(1) * (a_row.squeeze(1) != 0)
        d_alpha = torch.stack([s_row[tp[t]:tp[t + 1]].sum() for t in range(len(tp) - 1)])
        return d_lin, d_x, d_alpha * alpha_t * (1 - alpha_t), None, None


# ------------------------------------------------------------------------------------------------------------ HGT
class RelTransformFn(torch.autograd.Function):
    """y[y_idx[i]] = W[r(i), h] (.) x[x_idx[i]] for the relation-grouped (dst, relation) segments i (models/HGT.py:88-93 moved
    to the segments, see models/hgt.py).  x_idx may repeat (the q side gathers the dst row of every segment), y_idx and
    `order` are permutations of the segments.  Backward: the transposed maps through the same kernel; the weight gradient
    is a per-relation batched outer-product reduction (cuBLAS bmm, like the typed wgrad)."""

    @staticmethod
    def forward(ctx, x, w, x_idx, y_idx, rel_ptr_c, rel_ptr, R: int, H: int, d_k: int, w_kn: bool, n_out_rows: int,
                row_seg_ptr):
        x = x.contiguous()
        ctx.save_for_backward(x, w, x_idx, y_idx, row_seg_ptr if row_seg_ptr is not None else x_idx)
        ctx.meta = (rel_ptr_c, rel_ptr, R, H, d_k, w_kn, row_seg_ptr is not None)
        return ops.rel_transform(x, x_idx, y_idx, w.detach().contiguous(), rel_ptr_c, R, H, d_k, w_kn, n_out_rows)

    @staticmethod
    def backward(ctx, dy):
        x, w, x_idx, y_idx, row_seg_ptr = ctx.saved_tensors
        rel_ptr_c, rel_ptr, R, H, d_k, w_kn, reduce_rows = ctx.meta
        dy = dy.contiguous()
        S = int(y_idx.numel())
        dx = dw = None
        if ctx.needs_input_grad[0]:
            # d x_i = W^T-side map of d y_i, written in natural segment order, then summed over the segments that share
            # an input row (q side) or scattered back through the permutation (message side)
            dseg = ops.rel_transform(dy, y_idx, y_idx, w.detach().contiguous(), rel_ptr_c, R, H, d_k, not w_kn, S)
            if reduce_rows:
                ones = torch.ones(x.shape[0], dtype=torch.float32, device=x.device)
                dx = ops.segment_combine(dseg, row_seg_ptr, ones, x.shape[0], H * d_k)
            else:
                dx = torch.zeros_like(x)
                dx.index_copy_(0, x_idx.to(torch.int64), dseg.index_select(0, y_idx.to(torch.int64)))
        if ctx.needs_input_grad[1]:
            dw = torch.zeros_like(w)
            xi, yi = x_idx.to(torch.int64), y_idx.to(torch.int64)
            for r in range(R):
                a, b = rel_ptr[r], rel_ptr[r + 1]
                if b <= a:
                    continue
                xr = x.index_select(0, xi[a:b]).view(b - a, H, d_k).permute(1, 0, 2)          # [H, n, dk]
                gr = dy.index_select(0, yi[a:b]).view(b - a, H, d_k).permute(1, 0, 2)
                # w_kn: y_n = sum_k x_k W[k, n] -> dW[k, n] = sum_i x_k g_n;   else y_n = sum_k W[n, k] x_k -> dW[n, k] = g_n x_k
                dw[r] = torch.bmm(xr.transpose(1, 2), gr) if w_kn else torch.bmm(gr.transpose(1, 2), xr)
        return dx, dw, None, None, None, None, None, None, None, None, None, None


class SegAttnFn(torch.autograd.Function):
    """Edge attention over the (dst, relation) segments of HGT (models/HGT.py:95-106 before the cross-relation mean):
    out[s] = softmax_{e in s}(<q_s, k[src e]> / sqrt(d_k)) . v[src e]  - the HEAT kernels (forward K2, backward K3) run on
    the SEGMENT graph: one "row" per segment, one relation run per row, unit edge attribute; relation_pri is folded into
    q_s by the caller.  k, v, q_s in the lane-grouped column order."""

    @staticmethod
    def forward(ctx, k, v, qseg, plan, D: int, H: int):
        segs = plan.segments()
        aux = plan.cache.get("hgt_seg_graph")
        if aux is None:
            dev = qseg.device
            aux = dict(sim=torch.ones(plan.E, dtype=torch.float32, device=dev),
                       rel=torch.zeros(plan.E, dtype=torch.uint8, device=dev),
                       inv=torch.ones(segs["S"], dtype=torch.float32, device=dev),
                       ew=torch.ones(1, dtype=torch.float32, device=dev), eb=torch.zeros(1, dtype=torch.float32, device=dev))
            plan.cache["hgt_seg_graph"] = aux
        ctx.save_for_backward(k, v, qseg)
        ctx.plan, ctx.D, ctx.H, ctx.aux = plan, D, H, aux
        return ops.hetero_attn(k, v, qseg, segs["seg_ptr"], plan.e_src, aux["sim"], aux["rel"], aux["inv"], aux["ew"],
                               aux["eb"], D, H, True)

    @staticmethod
    def backward(ctx, d_out):
        k, v, qseg = ctx.saved_tensors
        plan, D, H, aux = ctx.plan, ctx.D, ctx.H, ctx.aux
        segs = plan.segments()
        if "transposed" not in aux:
            aux["transposed"] = ops.transposed_edges(segs["seg_ptr"], plan.e_src, plan.N)
        dk, dv = torch.empty((plan.N, D), dtype=torch.float32, device=k.device), torch.empty((plan.N, D), dtype=torch.float32, device=k.device)
        dq = torch.empty_like(qseg)
        ops.hetero_attn_bwd(k, v, qseg, segs["seg_ptr"], plan.e_src, aux["sim"], aux["rel"], aux["inv"], aux["ew"], aux["eb"],
                            D, H, d_out.contiguous(), dk, dv, dq, transposed=aux["transposed"])
        return dk, dv, dq, None, None, None


class SegmentCombineFn(torch.autograd.Function):
    """agg[v] = inv_r[v] * sum of the messages of row v's segments (the stack -> mean of models/HGT.py:105-106)."""

    @staticmethod
    def forward(ctx, msg, plan, D: int):
        segs = plan.segments()
        ctx.plan = plan
        return ops.segment_combine(msg.contiguous(), segs["row_seg_ptr"], plan.node_inv_r, plan.N, D)

    @staticmethod
    def backward(ctx, d_agg):
        plan = ctx.plan
        segs = plan.segments()
        dst = segs["seg_dst"].to(torch.int64)
        return d_agg.index_select(0, dst) * plan.node_inv_r.index_select(0, dst).unsqueeze(1), None, None
