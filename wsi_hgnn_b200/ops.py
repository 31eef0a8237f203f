"""Tensor-level wrappers of the C-ABI kernels (include/wsi_hgnn.h).

Every function validates dtype / device / strides, takes raw device pointers of torch tensors and
enqueues the kernel on torch's current CUDA stream (so torch.cuda.graph capture works).  PyTorch is
used for device memory and streams only; there is no CPU or PyTorch fallback - a CPU tensor raises.
"""
import ctypes
from functools import lru_cache
from typing import Optional, Sequence

import torch

from . import _lib

ACT_NONE, ACT_GELU = 0, 1
POOL_OPS = {"sum": 0, "mean": 1, "max": 2}
IMPL_AUTO, IMPL_SIMT, IMPL_TC = 0, 1, 2

# Operand formats of the tensor-core typed linear (WSI_OPF_*, include/wsi_hgnn.h) and the process-wide default
# (the analogue of torch.backends.cuda.matmul.allow_tf32 - but measured, see profiles/r2_precision_study.json):
#   "fp16"    one fp16 pass, fp32 accumulate: 11-bit significand = TF32's; the default of the fp32 models
#   "bf16x3"  3-term bf16 split, ~2^-17 relative: "exact" mode; always used for gradient GEMMs
#   "bf16"    one bf16 pass: the bf16-storage configuration (BASELINE config 3)
OPF_BF16X3, OPF_F16, OPF_BF16 = 0, 1, 2
_PRECISIONS = {"bf16x3": OPF_BF16X3, "fp16": OPF_F16, "bf16": OPF_BF16}
_OPF_DTYPE = {OPF_BF16X3: torch.bfloat16, OPF_F16: torch.float16, OPF_BF16: torch.bfloat16}
_opf_default = [OPF_F16]


def set_matmul_precision(name: str) -> str:
    """Operand precision of the tensor-core GEMMs of every later forward: "fp16" (default) | "bf16x3" | "bf16".
    Returns the previous setting."""
    if name not in _PRECISIONS:
        raise ValueError(f"matmul precision must be one of {sorted(_PRECISIONS)}, got {name!r}")
    prev = get_matmul_precision()
    _opf_default[0] = _PRECISIONS[name]
    return prev


def get_matmul_precision() -> str:
    return next(k for k, v in _PRECISIONS.items() if v == _opf_default[0])


class matmul_precision:
    """with ops.matmul_precision("bf16x3"): ...   (restores the previous setting on exit)"""

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        self.prev = set_matmul_precision(self.name)
        return self

    def __exit__(self, *exc):
        set_matmul_precision(self.prev)


def matmul_opf(opf: Optional[int] = None) -> int:
    return _opf_default[0] if opf is None else int(opf)


# The differentiable (training) path keeps the 3-term split for its forward GEMMs too: with an fp16-rounded forward the
# gradients of ill-conditioned parameters (e_linear.weight: a sum over all edges with heavy cancellation) moved by
# 2e-2 relative against the fp64 oracle, outside the 2e-3 gradient-parity bar; set_train_matmul_precision("fp16")
# trades that for speed.  Gradient GEMMs (dgrad / wgrad) always use the split.
_train_opf = [OPF_BF16X3]


def set_train_matmul_precision(name: str) -> str:
    if name not in _PRECISIONS:
        raise ValueError(f"matmul precision must be one of {sorted(_PRECISIONS)}, got {name!r}")
    prev = next(k for k, v in _PRECISIONS.items() if v == _train_opf[0])
    _train_opf[0] = _PRECISIONS[name]
    return prev


def train_opf() -> int:
    return _train_opf[0]


def operand_rows(rows: int, opf: int) -> int:
    return 2 * rows if opf == OPF_BF16X3 else rows


def dev_set(key: str, value: int):
    """Development knob of the library (wsi_dev_set); not product API."""
    _lib.check(_lib.load().wsi_dev_set(key.encode(), int(value)), "wsi_dev_set")

import threading

_cur_device = threading.local()      # cudaSetDevice is per host thread (the streaming evaluator plans on a worker thread)


def _prep(t: torch.Tensor):
    """The library has its own (static) CUDA runtime: keep its current device in step with the tensor's."""
    if not t.is_cuda:
        raise RuntimeError("wsi_hgnn_b200 ops need CUDA tensors: the hot path has no CPU fallback "
                           "(the CPU oracle under oracle/ is test infrastructure only)")
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if getattr(_cur_device, "idx", None) != idx:
        _lib.check(_lib.load().wsi_set_device(idx), "wsi_set_device")
        _cur_device.idx = idx
    return torch._C._cuda_getCurrentRawStream(idx)       # raw cudaStream_t of torch's current stream (no Stream object)


def _rows(t: Optional[torch.Tensor], name: str, dtype=torch.float32):
    """(pointer, row stride) of a 2-D row-strided tensor (unit column stride), or (None, 0)."""
    if t is None:
        return None, 0
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor")
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise ValueError(f"{name}: expected a 2-D tensor with unit column stride, got shape {tuple(t.shape)} "
                         f"strides {t.stride()}")
    return t.data_ptr(), (t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1]))


def _vec(t: Optional[torch.Tensor], name: str, dtype=torch.float32):
    if t is None:
        return None
    if t.dtype != dtype or not t.is_cuda or not t.is_contiguous():
        raise TypeError(f"{name}: expected a contiguous CUDA {dtype} tensor, got {t.dtype} on {t.device}")
    return t.data_ptr()


_KV_DTYPE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}
_DT_SIZE = {torch.int32: 4, torch.float32: 4, torch.uint8: 1, torch.int64: 8, torch.bfloat16: 2, torch.float16: 2}


def _arena(dev, specs):
    """Several device arrays out of ONE allocation (each torch.empty costs the host ~5 us: the planner of the streamed
    end-to-end path makes ~20 of them per slide).  specs: [(numel, dtype)] -> list of 1-D views, 256 B aligned."""
    offs, total = [], 0
    for n, dt in specs:
        offs.append(total)
        total += (max(int(n), 0) * _DT_SIZE[dt] + 255) // 256 * 256
    buf = torch.empty(max(total, 256), dtype=torch.uint8, device=dev)
    return [buf[o:o + max(int(n), 0) * _DT_SIZE[dt]].view(dt) for o, (n, dt) in zip(offs, specs)], buf


def host_i32(values: Sequence[int]):
    return (ctypes.c_int32 * len(values))(*[int(v) for v in values])


@lru_cache(maxsize=None)
def head_perm(D: int, H: int) -> Optional[torch.Tensor]:
    """Lane-grouped column order of the vector attention kernel, or None when (D, H) has none.
    perm[p] = logical column stored at physical position p."""
    if D % 128 != 0 or D > 1024 or H < 1 or H > 32 or (H & (H - 1)) != 0 or D % H != 0:
        return None
    buf = (ctypes.c_int32 * D)()
    _lib.check(_lib.load().wsi_head_perm(D, H, buf), "wsi_head_perm")
    return torch.tensor(list(buf), dtype=torch.int64)


def typed_linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], type_ptr: Sequence[int], *,
                 act: int = ACT_NONE, skip: Optional[torch.Tensor] = None, res: Optional[torch.Tensor] = None,
                 drop_mask: Optional[torch.Tensor] = None, row_gate: Optional[torch.Tensor] = None,
                 row_scale: Optional[torch.Tensor] = None, impl: int = IMPL_AUTO,
                 out: Optional[torch.Tensor] = None, type_ptr_c=None, opf: Optional[int] = None) -> torch.Tensor:
    """y[rows of type t] = epilogue(x[rows of type t] @ w[t].T); see wsi_typed_linear_f32.
    opf: operand format of the tensor-core path (None = the process default, set_matmul_precision)."""
    lib = _lib.load()
    opf = matmul_opf(opf)
    stream = _prep(x)
    T = len(type_ptr) - 1
    if w.dim() != 3 or w.shape[0] != T:
        raise ValueError(f"typed_linear: w must be [T={T}, n_out, K], got {tuple(w.shape)}")
    n_out, K = int(w.shape[1]), int(w.shape[2])
    N = int(type_ptr[-1])
    if x.shape[0] != N or x.shape[1] != K:
        raise ValueError(f"typed_linear: x is {tuple(x.shape)}, expected [{N}, {K}]")
    xp, ldx = _rows(x, "x")
    wp = _vec(w, "w")
    bp = _vec(bias, "bias")
    if bias is not None and tuple(bias.shape) != (T, n_out):
        raise ValueError("typed_linear: bias must be [T, n_out]")
    if out is None:
        out = torch.empty((N, n_out), dtype=torch.float32, device=x.device)
    yp, ldy = _rows(out, "out")
    rp, ldres = _rows(res, "res")
    mp, ldm = _rows(drop_mask, "drop_mask")
    ws_bytes = lib.wsi_typed_linear_workspace_bytes(N, K, n_out, T, impl, opf)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device) if ws_bytes > 0 else None
    tp = type_ptr_c if type_ptr_c is not None else host_i32(type_ptr)
    rc = lib.wsi_typed_linear_f32(xp, ldx, wp, bp, K, n_out, tp, T, act, _vec(skip, "skip"), rp, ldres, mp, ldm,
                                  _vec(row_gate, "row_gate"), _vec(row_scale, "row_scale"), yp, ldy, impl, opf,
                                  ws.data_ptr() if ws is not None else None, ws_bytes, stream)
    _lib.check(rc, "wsi_typed_linear_f32")
    return out


def tc_ok(n_rows: int, K: int, n_out: int) -> bool:
    """True when the tcgen05 typed linear takes this shape (wsi_typed_linear_tc_ok)."""
    return bool(_lib.load().wsi_typed_linear_tc_ok(int(n_rows), int(K), int(n_out)))


def to_operand(x: torch.Tensor, opf: Optional[int] = None) -> torch.Tensor:
    """fp32 [rows, K] -> the operand form of typed_linear_op (wsi_to_operand): bf16 [2 * rows, K] = [hi; lo] with
    x = hi + lo (OPF_BF16X3), fp16 [rows, K] (OPF_F16) or bf16 [rows, K] (OPF_BF16).  A weight stack [T, n_out, K] is
    converted as [T * n_out, K].  opf None = the process default (set_matmul_precision)."""
    lib = _lib.load()
    stream = _prep(x)
    opf = matmul_opf(opf)
    if x.dim() == 3:
        x = x.reshape(-1, x.shape[-1])
    xp, ld = _rows(x, "x")
    rows, K = int(x.shape[0]), int(x.shape[1])
    out = torch.empty((operand_rows(rows, opf), K), dtype=_OPF_DTYPE[opf], device=x.device)
    _lib.check(lib.wsi_to_operand(xp, ld, rows, K, opf, out.data_ptr(), stream), "wsi_to_operand")
    return out


def gather_to_operand(x: torch.Tensor, row_idx: torch.Tensor, opf: Optional[int] = None) -> torch.Tensor:
    """out row i = operand form of x[row_idx[i]] (wsi_gather_to_operand): fp32 [*, K] -> [operand_rows(len(row_idx)), K]."""
    lib = _lib.load()
    stream = _prep(x)
    opf = matmul_opf(opf)
    xp, ld = _rows(x, "x")
    rows, K = int(row_idx.numel()), int(x.shape[1])
    out = torch.empty((operand_rows(rows, opf), K), dtype=_OPF_DTYPE[opf], device=x.device)
    _lib.check(lib.wsi_gather_to_operand(xp, ld, _vec(row_idx, "row_idx", torch.int32), rows, K, opf, out.data_ptr(), stream),
               "wsi_gather_to_operand")
    return out


def to_operand_colsum(x: torch.Tensor, type_ptr: Sequence[int], type_ptr_c=None):
    """fp32 [N, K] -> (OPF_BF16X3 operand form [2N, K], per-type column sums [T, K]) in one pass; see wsi_to_operand_colsum."""
    lib = _lib.load()
    stream = _prep(x)
    xp, ld = _rows(x, "x")
    N, K = int(x.shape[0]), int(x.shape[1])
    T = len(type_ptr) - 1
    if int(type_ptr[-1]) != N:
        raise ValueError("to_operand_colsum: type_ptr does not cover the rows of x")
    tpc = type_ptr_c if type_ptr_c is not None else host_i32(type_ptr)
    out = torch.empty((2 * N, K), dtype=torch.bfloat16, device=x.device)
    cs = torch.empty((T, K), dtype=torch.float32, device=x.device)
    ws_bytes = int(lib.wsi_to_operand_colsum_workspace_bytes(K, tpc, T))
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=x.device)
    _lib.check(lib.wsi_to_operand_colsum(xp, ld, K, tpc, T, out.data_ptr(), cs.data_ptr(), ws.data_ptr(), ws_bytes, stream),
               "wsi_to_operand_colsum")
    return out, cs


def gather_rows16(x_op: torch.Tensor, row_idx: torch.Tensor) -> torch.Tensor:
    """out row i = x_op[row_idx[i]] for a single-plane 16-bit operand matrix (OPF_F16 / OPF_BF16); see wsi_gather_rows16."""
    lib = _lib.load()
    stream = _prep(x_op)
    if x_op.dtype not in (torch.float16, torch.bfloat16):
        raise ValueError("gather_rows16: expected an fp16 / bf16 matrix")
    xp, ld = _rows(x_op, "x_op", x_op.dtype)              # (a column slice of a wider matrix is fine)
    rows, K = int(row_idx.numel()), int(x_op.shape[1])
    out = torch.empty((rows, K), dtype=x_op.dtype, device=x_op.device)
    _lib.check(lib.wsi_gather_rows16(xp, ld, _vec(row_idx, "row_idx", torch.int32), rows, K, out.data_ptr(), stream),
               "wsi_gather_rows16")
    return out


def typed_linear_op(x_op: torch.Tensor, w_op: torch.Tensor, bias: Optional[torch.Tensor],
                    type_ptr: Sequence[int], n_out: int, *, act: int = ACT_NONE,
                    skip: Optional[torch.Tensor] = None, res: Optional[torch.Tensor] = None,
                    drop_mask: Optional[torch.Tensor] = None, row_gate: Optional[torch.Tensor] = None,
                    row_scale: Optional[torch.Tensor] = None, want_y: bool = True, want_op: bool = False,
                    type_ptr_c=None, opf: Optional[int] = None, out: Optional[torch.Tensor] = None,
                    out_op: Optional[torch.Tensor] = None):
    """tcgen05 typed linear on operands already in operand form (to_operand); see wsi_typed_linear_op.
    -> y fp32 [N, n_out] (or None), y in operand form (or None).  out / out_op: caller-owned destinations (out_op: a
    dense, contiguous tensor of the operand form's shape and dtype, e.g. a rank's block of an all-gather buffer)."""
    lib = _lib.load()
    stream = _prep(x_op)
    opf = matmul_opf(opf)
    T = len(type_ptr) - 1
    N, K = int(type_ptr[-1]), int(x_op.shape[1])
    dt = _OPF_DTYPE[opf]
    for name, t, rows in (("x_op", x_op, operand_rows(N, opf)), ("w_op", w_op, operand_rows(T * n_out, opf))):
        if t.dtype != dt or not t.is_contiguous() or tuple(t.shape) != (rows, K):
            raise ValueError(f"typed_linear_op: {name} must be a contiguous {dt} [{rows}, {K}] tensor, "
                             f"got {t.dtype} {tuple(t.shape)}")
    y = (out if out is not None else torch.empty((N, n_out), dtype=torch.float32, device=x_op.device)) if want_y else None
    if y is not None and (tuple(y.shape) != (N, n_out) or y.dtype != torch.float32 or not y.is_contiguous()):
        raise ValueError("typed_linear_op: out must be a contiguous fp32 [N, n_out] tensor")
    ys = (out_op if out_op is not None else torch.empty((operand_rows(N, opf), n_out), dtype=dt, device=x_op.device)) if want_op else None
    if ys is not None and (tuple(ys.shape) != (operand_rows(N, opf), n_out) or ys.dtype != dt or not ys.is_contiguous()):
        raise ValueError(f"typed_linear_op: out_op must be a contiguous {dt} [{operand_rows(N, opf)}, {n_out}] tensor")
    rp, ldres = _rows(res, "res")
    mp, ldm = _rows(drop_mask, "drop_mask")
    tp = type_ptr_c if type_ptr_c is not None else host_i32(type_ptr)
    rc = lib.wsi_typed_linear_op(x_op.data_ptr(), w_op.data_ptr(), _vec(bias, "bias"), K, n_out, tp, T, act,
                                 _vec(skip, "skip"), rp, ldres, mp, ldm, _vec(row_gate, "row_gate"),
                                 _vec(row_scale, "row_scale"), y.data_ptr() if y is not None else None, n_out,
                                 ys.data_ptr() if ys is not None else None, opf, stream)
    _lib.check(rc, "wsi_typed_linear_op")
    return y, ys


def hetero_attn(k: torch.Tensor, v: torch.Tensor, q: torch.Tensor, rowptr: torch.Tensor, e_src: torch.Tensor,
                e_sim: torch.Tensor, e_rel: torch.Tensor, node_inv_r: torch.Tensor, e_w: torch.Tensor,
                e_b: torch.Tensor, D: int, H: int, use_head_perm: bool, want_attn: bool = False):
    """HEAT edge attention for all relations of one layer; see wsi_hetero_attn_fwd."""
    lib = _lib.load()
    stream = _prep(q)
    N = int(q.shape[0])
    kp, ldk = _rows(k, "k")
    vp, ldv = _rows(v, "v")
    qp, ldq = _rows(q, "q")
    agg = torch.empty((N, D), dtype=torch.float32, device=q.device)
    attn = torch.empty((int(e_src.shape[0]), H), dtype=torch.float32, device=q.device) if want_attn else None
    rc = lib.wsi_hetero_attn_fwd(kp, ldk, vp, ldv, qp, ldq, _vec(rowptr, "rowptr", torch.int32),
                                 _vec(e_src, "e_src", torch.int32), _vec(e_sim, "e_sim"),
                                 _vec(e_rel, "e_rel", torch.uint8), _vec(node_inv_r, "node_inv_r"),
                                 _vec(e_w.reshape(-1), "e_w"), _vec(e_b.reshape(-1), "e_b"), N, D, H,
                                 1 if use_head_perm else 0, agg.data_ptr(), D,
                                 attn.data_ptr() if attn is not None else None, stream)
    _lib.check(rc, "wsi_hetero_attn_fwd")
    return (agg, attn) if want_attn else agg


def hetero_attn_work(k: torch.Tensor, v: torch.Tensor, q: torch.Tensor, work: dict, e_src: torch.Tensor,
                     e_sim: torch.Tensor, e_rel: torch.Tensor, node_inv_r: torch.Tensor, e_w: torch.Tensor,
                     e_b: torch.Tensor, D: int, H: int, out: Optional[torch.Tensor] = None,
                     op_out: bool = False, opf: Optional[int] = None, n_rows: Optional[int] = None) -> torch.Tensor:
    """HEAT edge attention driven by the hub-balancing work list of GraphPlan.attn_work();
    see wsi_hetero_attn_work_fwd.  k/v/q/agg columns are in the head_perm(D, H) order.
    op_out: return the result in operand form (to_operand layout, the A operand of typed_linear_op) instead of fp32.
    q may be fp32 / fp16 / bf16; its row count is the number of output rows (HGT: segments), k / v rows are gathered."""
    lib = _lib.load()
    stream = _prep(q)
    opf = matmul_opf(opf)
    N = int(q.shape[0]) if n_rows is None else int(n_rows)
    if q.dtype not in _KV_DTYPE:
        raise TypeError(f"hetero_attn_work: q must be fp32, fp16 or bf16, got {q.dtype}")
    if k.dtype != v.dtype or k.dtype not in _KV_DTYPE:
        raise TypeError(f"hetero_attn_work: k / v must both be fp32, fp16 or bf16, got {k.dtype} / {v.dtype}")
    kp, ldk = _rows(k, "k", k.dtype)
    vp, ldv = _rows(v, "v", v.dtype)
    qp, ldq = _rows(q, "q", q.dtype)
    agg = agg_split = None
    if op_out:
        agg_split = torch.empty((operand_rows(N, opf), D), dtype=_OPF_DTYPE[opf], device=q.device)
        ap, ldo = None, D
    else:
        agg = out if out is not None else torch.empty((N, D), dtype=torch.float32, device=q.device)
        ap, ldo = _rows(agg, "agg")
    n_part, n_split = work["n_part"], work["n_split"]
    part_ms = part_acc = None
    if n_part > 0:
        part_ms = torch.empty((n_part, 64), dtype=torch.float32, device=q.device)
        part_acc = torch.empty((n_part, D), dtype=torch.float32, device=q.device)
    rc = lib.wsi_hetero_attn_work_fwd(kp, ldk, vp, ldv, _KV_DTYPE[k.dtype], qp, _KV_DTYPE[q.dtype], ldq,
                                      _vec(e_src, "e_src", torch.int32), _vec(e_sim, "e_sim"),
                                      _vec(e_rel, "e_rel", torch.uint8), _vec(node_inv_r, "node_inv_r"),
                                      _vec(e_w.reshape(-1), "e_w"), _vec(e_b.reshape(-1), "e_b"), N, int(k.shape[0]), D, H,
                                      _vec(work["items"], "items", torch.int32), work["n_items"],
                                      _vec(work["split_row"], "split_row", torch.int32),
                                      _vec(work["split_ptr"], "split_ptr", torch.int32),
                                      _vec(work["part_rel"], "part_rel", torch.int32),
                                      _vec(work.get("part_split"), "part_split", torch.int32),
                                      _vec(work.get("split_cnt"), "split_cnt", torch.int32),
                                      _vec(work.get("sched"), "sched", torch.int32), n_split, n_part,
                                      part_ms.data_ptr() if part_ms is not None else None,
                                      part_acc.data_ptr() if part_acc is not None else None, ap, ldo,
                                      agg_split.data_ptr() if agg_split is not None else None, opf, stream)
    _lib.check(rc, "wsi_hetero_attn_work_fwd")
    return agg_split if op_out else agg


def transposed_edges(rowptr: torch.Tensor, e_src: torch.Tensor, n_src: int):
    """Source-major view of a dst-major CSR: (t_ptr int32 [n_src + 1], t_eid int32 [E] = position of the edge in the
    dst-major arrays, t_dst int32 [E] = its dst row), out-edges of a source in dst-major order (deterministic)."""
    dev = rowptr.device
    n_rows = int(rowptr.numel()) - 1
    deg = (rowptr[1:] - rowptr[:-1]).to(torch.int64)
    e_dst = torch.repeat_interleave(torch.arange(n_rows, device=dev, dtype=torch.int64), deg)
    src64 = e_src.to(torch.int64)
    order = torch.argsort(src64, stable=True)
    t_ptr = torch.zeros(n_src + 1, dtype=torch.int32, device=dev)
    if src64.numel():
        t_ptr[1:] = torch.cumsum(torch.bincount(src64, minlength=n_src), 0).to(torch.int32)
    return t_ptr, order.to(torch.int32).contiguous(), e_dst[order].to(torch.int32).contiguous()


def hetero_attn_bwd(k, v, q, rowptr, e_src, e_sim, e_rel, node_inv_r, e_w, e_b, D: int, H: int, d_agg: torch.Tensor,
                    dk: torch.Tensor, dv: torch.Tensor, dq: torch.Tensor,
                    row_order: Optional[torch.Tensor] = None, transposed=None) -> torch.Tensor:
    """Backward of the HEAT edge attention; see wsi_hetero_attn_bwd.  dq is written.  Without `transposed`: dk / dv must
    be zero-filled (vector atomics accumulate into them).  With transposed = transposed_edges(rowptr, e_src, n_src):
    the two-pass mode without atomics, dk / dv (all n_src rows) are written.
    row_order: int32 [N] processing order of the rows (GraphPlan.rows_by_degree()).
    -> d_e fp32 [2] = (d e_linear.weight, d e_linear.bias)."""
    lib = _lib.load()
    stream = _prep(q)
    N = int(q.shape[0])
    kp, ldk = _rows(k, "k")
    vp, ldv = _rows(v, "v")
    qp, ldq = _rows(q, "q")
    gp, ldg = _rows(d_agg, "d_agg")
    dkp, lddk = _rows(dk, "dk")
    dvp, lddv = _rows(dv, "dv")
    dqp, lddq = _rows(dq, "dq")
    d_e = torch.zeros(2, dtype=torch.float32, device=q.device)
    t_ptr = t_eid = t_dst = coef = None
    n_src = 0
    if transposed is not None:
        t_ptr, t_eid, t_dst = transposed
        n_src = int(t_ptr.numel()) - 1
        if dk.shape[0] != n_src or dv.shape[0] != n_src:
            raise ValueError("hetero_attn_bwd: dk / dv must have one row per source row of the transposed edge list")
        coef = torch.zeros((int(e_src.numel()), 2 * H), dtype=torch.float32, device=q.device)
    rc = lib.wsi_hetero_attn_bwd(kp, ldk, vp, ldv, qp, ldq, _vec(rowptr, "rowptr", torch.int32),
                                 _vec(e_src, "e_src", torch.int32), _vec(e_sim, "e_sim"),
                                 _vec(e_rel, "e_rel", torch.uint8), _vec(node_inv_r, "node_inv_r"),
                                 _vec(e_w.reshape(-1), "e_w"), _vec(e_b.reshape(-1), "e_b"), N, D, H, gp, ldg, dkp, lddk,
                                 dvp, lddv, dqp, lddq, d_e.data_ptr(), _vec(row_order, "row_order", torch.int32),
                                 _vec(t_ptr, "t_ptr", torch.int32), _vec(t_eid, "t_eid", torch.int32),
                                 _vec(t_dst, "t_dst", torch.int32), n_src, coef.data_ptr() if coef is not None else None,
                                 stream)
    _lib.check(rc, "wsi_hetero_attn_bwd")
    return d_e


def hetero_attn_seg(k, v, qseg, seg_ptr, seg_rel, e_src, rel_pri, D: int, H: int, use_head_perm: bool = False,
                    items: Optional[torch.Tensor] = None, want_out: bool = True, op_out: bool = False,
                    opf: Optional[int] = None):
    """HGT edge attention over (dst, relation) segments; see wsi_hetero_attn_seg_fwd.
    items int32 [S, 4]: work list (row, e_beg, e_end, -1) for a segment order other than the edge order.
    -> out fp32 [S, D]; with op_out: (out or None, operand-form copy [operand_rows(S), D])."""
    lib = _lib.load()
    stream = _prep(qseg)
    S = int(qseg.shape[0])
    if k.dtype != v.dtype or k.dtype not in _KV_DTYPE:
        raise TypeError(f"hetero_attn_seg: k / v must both be fp32, fp16 or bf16, got {k.dtype} / {v.dtype}")
    kp, ldk = _rows(k, "k", k.dtype)
    vp, ldv = _rows(v, "v", v.dtype)
    if qseg.dtype not in _KV_DTYPE:
        raise TypeError(f"hetero_attn_seg: qseg must be fp32, fp16 or bf16, got {qseg.dtype}")
    qp, ldq = _rows(qseg, "qseg", qseg.dtype)
    opf = matmul_opf(opf)
    out = torch.empty((S, D), dtype=torch.float32, device=qseg.device) if want_out or not op_out else None
    out_op = torch.empty((operand_rows(S, opf), D), dtype=_OPF_DTYPE[opf], device=qseg.device) if op_out else None
    if items is not None and (items.dtype != torch.int32 or tuple(items.shape) != (S, 4) or not items.is_contiguous()):
        raise ValueError(f"hetero_attn_seg: items must be a contiguous int32 [{S}, 4] tensor")
    rc = lib.wsi_hetero_attn_seg_fwd(kp, ldk, vp, ldv, _KV_DTYPE[k.dtype], qp, _KV_DTYPE[qseg.dtype], ldq,
                                     _vec(seg_ptr, "seg_ptr", torch.int32) if seg_ptr is not None else None,
                                     _vec(seg_rel, "seg_rel", torch.int32), _vec(e_src, "e_src", torch.int32),
                                     _vec(rel_pri, "rel_pri"), S, D, H, 1 if use_head_perm else 0,
                                     items.data_ptr() if items is not None else None, int(k.shape[0]),
                                     out.data_ptr() if out is not None else None, D,
                                     out_op.data_ptr() if out_op is not None else None, opf, stream)
    _lib.check(rc, "wsi_hetero_attn_seg_fwd")
    return (out, out_op) if op_out else out


def rel_transform(x, x_row_idx, y_row_idx, w, rel_ptr_c, R: int, H: int, d_k: int, w_kn: bool, n_out_rows: int):
    """Per-(relation, head) d_k x d_k transform of relation-grouped segments; see wsi_rel_transform."""
    lib = _lib.load()
    stream = _prep(x)
    xp, ldx = _rows(x, "x")
    if tuple(w.shape) != (R, H, d_k, d_k):
        raise ValueError(f"rel_transform: w must be [{R}, {H}, {d_k}, {d_k}], got {tuple(w.shape)}")
    y = torch.empty((n_out_rows, H * d_k), dtype=torch.float32, device=x.device)
    rc = lib.wsi_rel_transform(xp, ldx, _vec(x_row_idx, "x_row_idx", torch.int32),
                               _vec(y_row_idx, "y_row_idx", torch.int32), _vec(w, "w"), rel_ptr_c, R, H, d_k,
                               1 if w_kn else 0, y.data_ptr(), H * d_k, stream)
    _lib.check(rc, "wsi_rel_transform")
    return y


def segment_combine(msg, row_seg_ptr, node_inv_r, N: int, D: int, seg_pos: Optional[torch.Tensor] = None,
                    want_out: bool = True, op_out: bool = False, opf: Optional[int] = None):
    """agg[v] = 1/R_v * sum of the messages of row v's (v, relation) segments; see wsi_segment_combine.
    seg_pos int32 [S]: msg row of segment s.  With op_out: (agg or None, operand-form copy)."""
    lib = _lib.load()
    stream = _prep(node_inv_r)
    if msg.dtype not in _KV_DTYPE:
        raise TypeError(f"segment_combine: msg must be fp32, fp16 or bf16, got {msg.dtype}")
    mp, ldm = _rows(msg, "msg", msg.dtype)
    opf = matmul_opf(opf)
    agg = torch.empty((N, D), dtype=torch.float32, device=node_inv_r.device) if want_out or not op_out else None
    agg_op = torch.empty((operand_rows(N, opf), D), dtype=_OPF_DTYPE[opf], device=node_inv_r.device) if op_out else None
    rc = lib.wsi_segment_combine(mp, _KV_DTYPE[msg.dtype], ldm, _vec(row_seg_ptr, "row_seg_ptr", torch.int32),
                                 _vec(seg_pos, "seg_pos", torch.int32) if seg_pos is not None else None,
                                 _vec(node_inv_r, "node_inv_r"), N, D, agg.data_ptr() if agg is not None else None, D,
                                 agg_op.data_ptr() if agg_op is not None else None, opf, stream)
    _lib.check(rc, "wsi_segment_combine")
    return (agg, agg_op) if op_out else agg


def typed_layernorm(x, gamma, beta, type_ptr: Sequence[int], eps: float = 1e-5, type_ptr_c=None, inplace=False,
                    row_gate: Optional[torch.Tensor] = None, op_out: bool = False, opf: Optional[int] = None):
    """Per-node-type LayerNorm (row-gated); with op_out -> (y, operand-form copy of y); see wsi_typed_layernorm."""
    lib = _lib.load()
    stream = _prep(x)
    T = len(type_ptr) - 1
    D = int(x.shape[1])
    if tuple(gamma.shape) != (T, D) or tuple(beta.shape) != (T, D):
        raise ValueError("typed_layernorm: gamma/beta must be [T, D]")
    xp, ldx = _rows(x, "x")
    y = x if inplace else torch.empty_like(x)
    yp, ldy = _rows(y, "y")
    tp = type_ptr_c if type_ptr_c is not None else host_i32(type_ptr)
    opf = matmul_opf(opf)
    y_op = torch.empty((operand_rows(int(x.shape[0]), opf), D), dtype=_OPF_DTYPE[opf], device=x.device) if op_out else None
    rc = lib.wsi_typed_layernorm(xp, ldx, _vec(gamma, "gamma"), _vec(beta, "beta"), _vec(row_gate, "row_gate"), tp, T, D,
                                 float(eps), yp, ldy, y_op.data_ptr() if y_op is not None else None, opf, stream)
    _lib.check(rc, "wsi_typed_layernorm")
    return (y, y_op) if op_out else y


def skip_mix_bwd(dout, out, x, drop_mask, skip, row_gate, type_ptr: Sequence[int], type_ptr_c=None):
    """Backward of the fused a_linear epilogue; see wsi_skip_mix_bwd.  -> d_lin [N, D], d_x [N, D], d_alpha [T]."""
    lib = _lib.load()
    stream = _prep(dout)
    T = len(type_ptr) - 1
    N, D = int(dout.shape[0]), int(dout.shape[1])
    gp, ldd = _rows(dout, "dout")
    op, ldo = _rows(out, "out")
    xp, ldx = _rows(x, "x")
    mp, ldm = _rows(drop_mask, "drop_mask")
    d_lin = torch.empty((N, D), dtype=torch.float32, device=dout.device)
    d_x = torch.empty((N, D), dtype=torch.float32, device=dout.device)
    d_alpha = torch.empty(T, dtype=torch.float32, device=dout.device)
    tp = type_ptr_c if type_ptr_c is not None else host_i32(type_ptr)
    rc = lib.wsi_skip_mix_bwd(gp, ldd, op, ldo, xp, ldx, mp, ldm, _vec(skip, "skip"), _vec(row_gate, "row_gate"), tp, T, D,
                              d_lin.data_ptr(), D, d_x.data_ptr(), D, d_alpha.data_ptr(), stream)
    _lib.check(rc, "wsi_skip_mix_bwd")
    return d_lin, d_x, d_alpha


_TYPE_PTR_DEV = {}


def type_ptr_dev(type_ptr: Sequence[int], device) -> torch.Tensor:
    """int32 device copy of a type_ptr list (cached: a pageable host -> device copy per call would synchronise)."""
    key = (tuple(int(v) for v in type_ptr), str(device))
    t = _TYPE_PTR_DEV.get(key)
    if t is None:
        if len(_TYPE_PTR_DEV) > 256:
            _TYPE_PTR_DEV.clear()
        t = torch.tensor(list(key[0]), dtype=torch.int32).to(device)
        _TYPE_PTR_DEV[key] = t
    return t


def typed_wgrad_ok(N: int, M: int, Nn: int, T: int) -> bool:
    return bool(_lib.load().wsi_typed_wgrad_supported(int(N), int(M), int(Nn), int(T)))


def typed_wgrad(dy_op: torch.Tensor, x_op: torch.Tensor, type_ptr: Sequence[int], type_ptr_c=None) -> torch.Tensor:
    """dw[t] = dY_t^T X_t  [T, M, Nn] fp32 on tcgen05 from the OPF_BF16X3 operand forms of dy ([2N, M]) and x ([2N, Nn]);
    see wsi_typed_wgrad."""
    lib = _lib.load()
    stream = _prep(dy_op)
    T = len(type_ptr) - 1
    N = int(type_ptr[-1])
    M, Nn = int(dy_op.shape[1]), int(x_op.shape[1])
    for name, t, cols in (("dy_op", dy_op, M), ("x_op", x_op, Nn)):
        if t.dtype != torch.bfloat16 or not t.is_contiguous() or tuple(t.shape) != (2 * N, cols):
            raise ValueError(f"typed_wgrad: {name} must be a contiguous bf16 [{2 * N}, {cols}] tensor, got {t.dtype} {tuple(t.shape)}")
    tpc = type_ptr_c if type_ptr_c is not None else host_i32(type_ptr)
    dw = torch.empty((T, M, Nn), dtype=torch.float32, device=dy_op.device)
    ws_bytes = int(lib.wsi_typed_wgrad_workspace_bytes(M, Nn, tpc, T))
    if ws_bytes < 0:
        raise ValueError("typed_wgrad: bad type_ptr")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dy_op.device)
    _lib.check(lib.wsi_typed_wgrad(dy_op.data_ptr(), x_op.data_ptr(), M, Nn, tpc, T, dw.data_ptr(), ws.data_ptr(), ws_bytes,
                                   stream), "wsi_typed_wgrad")
    return dw


def typed_colsum(dy: torch.Tensor, type_ptr: Sequence[int]) -> torch.Tensor:
    """[T, n_out] column sums of dy over the rows of every type (the bias gradient of a typed linear): the typed readout
    kernel with the types as segments."""
    return segment_pool(dy, type_ptr_dev(type_ptr, dy.device), len(type_ptr) - 1, "sum")


def segment_pool(x: torch.Tensor, seg_ptr: torch.Tensor, n_seg: int, op: str) -> torch.Tensor:
    """[n_seg, D] typed readout over (type, graph) row segments; see wsi_segment_pool_fwd."""
    if op not in POOL_OPS:
        raise NotImplementedError(op)
    lib = _lib.load()
    stream = _prep(x)
    N, D = int(x.shape[0]), int(x.shape[1])
    xp, ldx = _rows(x, "x")
    out = torch.empty((n_seg, D), dtype=torch.float32, device=x.device)
    ws_bytes = lib.wsi_segment_pool_workspace_bytes(N, n_seg, D)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device) if ws_bytes > 0 else None
    rc = lib.wsi_segment_pool_fwd(xp, ldx, _vec(seg_ptr, "seg_ptr", torch.int32), n_seg, N, D, POOL_OPS[op],
                                  out.data_ptr(), D, ws.data_ptr() if ws is not None else None, ws_bytes, stream)
    _lib.check(rc, "wsi_segment_pool_fwd")
    return out


def plan_build_csr(src: torch.Tensor, dst: torch.Tensor, sim: Optional[torch.Tensor], rel_table: torch.Tensor, R: int,
                   n_nodes: int, want_dst: bool = False):
    """Relation-grouped dst-major CSR of the packed graph; see wsi_plan_build_csr.
    -> rowptr, e_src, e_sim, e_rel, e_dst (or None), stats (device int32 [4])."""
    lib = _lib.load()
    stream = _prep(rel_table)
    dev = rel_table.device
    E = int(src.shape[0])
    ws_bytes = lib.wsi_plan_workspace_bytes(n_nodes, E)
    (rowptr, e_src, e_sim, e_rel, e_dst, stats, ws), _ = _arena(dev, [
        (n_nodes + 1, torch.int32), (E, torch.int32), (E, torch.float32), (E, torch.uint8),
        (E if want_dst else 0, torch.int32), (4, torch.int32), (ws_bytes, torch.uint8)])
    if not want_dst:
        e_dst = None
    sim32 = sim64 = None
    if sim is not None:
        if sim.dtype == torch.float64:
            sim64 = _vec(sim, "sim", torch.float64)
        else:
            sim32 = _vec(sim, "sim", torch.float32)
    rc = lib.wsi_plan_build_csr(_vec(src, "src", torch.int64), _vec(dst, "dst", torch.int64), sim32, sim64,
                                _vec(rel_table, "rel_table", torch.int32), R, n_nodes, E, rowptr.data_ptr(),
                                e_src.data_ptr(), e_sim.data_ptr(), e_rel.data_ptr(),
                                e_dst.data_ptr() if e_dst is not None else None, stats.data_ptr(), ws.data_ptr(),
                                ws_bytes, stream)
    _lib.check(rc, "wsi_plan_build_csr")
    return rowptr, e_src, e_sim, e_rel, e_dst, stats


def plan_attn_work_begin(rowptr: torch.Tensor, e_rel: torch.Tensor, n_nodes: int, chunk: int,
                         stats: Optional[torch.Tensor] = None) -> dict:
    """First half of plan_attn_work: the counting kernels and an ASYNCHRONOUS read of the totals into pinned host
    memory (event recorded on the current stream).  The caller can enqueue other work before plan_attn_work_finish
    waits for that event - the streaming evaluator launches the previous slide's forward in between."""
    lib = _lib.load()
    stream = _prep(rowptr)
    dev = rowptr.device
    ws_bytes = lib.wsi_plan_workspace_bytes(n_nodes, 0)
    (scans, hist, ws, totals), _ = _arena(dev, [(2 * (n_nodes + 1), torch.int32), (2 * (chunk + 1), torch.int32),
                                                (ws_bytes, torch.uint8), (4, torch.int32)])
    scans = scans.view(2, n_nodes + 1)
    rp, rl = _vec(rowptr, "rowptr", torch.int32), _vec(e_rel, "e_rel", torch.uint8)
    _lib.check(lib.wsi_plan_attn_work_count(rp, rl, n_nodes, chunk, scans[0].data_ptr(), scans[1].data_ptr(),
                                            hist.data_ptr(), ws.data_ptr(), ws_bytes, stream),
               "wsi_plan_attn_work_count")
    totals[:2] = scans[:, n_nodes]
    if stats is not None:
        totals[2:] = stats[:2]
    host = torch.empty(4, dtype=torch.int32, pin_memory=True)
    host.copy_(totals, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    return dict(rowptr=rowptr, e_rel=e_rel, n_nodes=n_nodes, chunk=chunk, scans=scans, hist=hist, ws=ws, host=host, ev=ev,
                has_stats=stats is not None)


def plan_attn_work_finish(ctx: dict) -> dict:
    """Second half: wait for the totals (the one host sync of the planner), size the arrays, fill them."""
    lib = _lib.load()
    rowptr, e_rel, n_nodes, chunk, scans, hist = (ctx[k] for k in ("rowptr", "e_rel", "n_nodes", "chunk", "scans", "hist"))
    stream = _prep(rowptr)
    dev = rowptr.device
    ctx["ev"].synchronize()
    n_part, n_split, max_deg, bad = ctx["host"].tolist()
    if not ctx["has_stats"]:
        max_deg, bad = None, 0
    n_items = n_part + n_nodes - n_split
    i32 = torch.int32
    # (split_cnt | sched) adjacent: the arrival counters of the fused merge (self-resetting) and the queue words, zeroed once
    (items, split_row, split_ptr, part_rel, part_split, zeroed), _ = _arena(dev, [
        (4 * max(n_items, 1), i32), (max(n_split, 1), i32), (n_split + 1, i32), (max(n_part, 1), i32),
        (max(n_part, 1), i32), (max(n_split, 1) + 64, i32)])
    items = items.view(max(n_items, 1), 4)
    zeroed.zero_()
    split_cnt, sched = zeroed[:max(n_split, 1)], zeroed[max(n_split, 1) + 62:max(n_split, 1) + 64]
    rp, rl = _vec(rowptr, "rowptr", torch.int32), _vec(e_rel, "e_rel", torch.uint8)
    _lib.check(lib.wsi_plan_attn_work_fill(rp, rl, n_nodes, chunk, scans[0].data_ptr(), scans[1].data_ptr(), n_part,
                                           n_split, hist.data_ptr(), items.data_ptr(), split_row.data_ptr(),
                                           split_ptr.data_ptr(), part_rel.data_ptr(), part_split.data_ptr(), stream),
               "wsi_plan_attn_work_fill")
    return dict(items=items, n_items=n_items, split_row=split_row, split_ptr=split_ptr, part_rel=part_rel,
                part_split=part_split, split_cnt=split_cnt, sched=sched,
                n_split=n_split, n_part=n_part, max_in_degree=max_deg, bad_edges=bool(bad))


def plan_attn_work(rowptr: torch.Tensor, e_rel: torch.Tensor, n_nodes: int, chunk: int,
                   stats: Optional[torch.Tensor] = None) -> dict:
    """Hub-balancing work list of hetero_attn_work; see wsi_plan_attn_work_count / _fill (one host sync for the
    two totals that size the arrays)."""
    return plan_attn_work_finish(plan_attn_work_begin(rowptr, e_rel, n_nodes, chunk, stats))


AFFINE_MAX_OUT = 8


def segment_pool_affine(x: torch.Tensor, seg_ptr: torch.Tensor, T: int, B: int, op: str, M: torch.Tensor,
                        c: Optional[torch.Tensor], b_total: Optional[torch.Tensor], seg_scale: Optional[torch.Tensor],
                        out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """Typed readout fused with the narrow affine prediction that follows it -> [B, n_out];
    see wsi_segment_pool_affine_fwd."""
    if op not in POOL_OPS:
        raise NotImplementedError(op)
    lib = _lib.load()
    stream = _prep(x)
    N, D = int(x.shape[0]), int(x.shape[1])
    n_out = int(M.shape[1])
    if tuple(M.shape) != (T, n_out, D) or n_out > AFFINE_MAX_OUT:
        raise ValueError(f"segment_pool_affine: M must be [T={T}, n_out<={AFFINE_MAX_OUT}, D={D}], got {tuple(M.shape)}")
    xp, ldx = _rows(x, "x")
    if out is None:
        out = torch.empty((B, n_out), dtype=torch.float32, device=x.device)
        accumulate = False
    ws_bytes = lib.wsi_segment_pool_affine_workspace_bytes(N, T * B, D)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=x.device)
    rc = lib.wsi_segment_pool_affine_fwd(xp, ldx, _vec(seg_ptr, "seg_ptr", torch.int32), T, B, N, D, POOL_OPS[op],
                                         _vec(M, "M"), _vec(c, "c"), _vec(b_total, "b_total"),
                                         _vec(seg_scale, "seg_scale"), n_out, 1 if accumulate else 0, out.data_ptr(),
                                         n_out, ws.data_ptr(), ws_bytes, stream)
    _lib.check(rc, "wsi_segment_pool_affine_fwd")
    return out


def knn_topk(feat: torch.Tensor, topn: int, q_begin: int = 0, q_end: Optional[int] = None, want_dist: bool = False):
    """Exact L2 k-NN (self included, ordered by (distance, index)); see wsi_knn_topk.
    -> int32 [q_end - q_begin, topn] (and the fp32 distances)."""
    lib = _lib.load()
    stream = _prep(feat)
    if feat.dtype != torch.float32 or feat.dim() != 2 or not feat.is_contiguous():
        raise TypeError("knn_topk: features must be a contiguous fp32 [N, F] tensor")
    n, F = int(feat.shape[0]), int(feat.shape[1])
    q_end = n if q_end is None else q_end
    if topn > n:
        # HNSW would return < radius hits -> np.stack fails -> ValueError (graph_constructor.py:268-272)
        raise ValueError(f"fewer than topn={topn} nodes (n={n})")
    nq = q_end - q_begin
    nbr = torch.empty((nq, topn), dtype=torch.int32, device=feat.device)
    dist = torch.empty((nq, topn), dtype=torch.float32, device=feat.device) if want_dist else None
    ws_bytes = lib.wsi_knn_workspace_bytes(n, F, topn, q_begin, q_end)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=feat.device)
    rc = lib.wsi_knn_topk(feat.data_ptr(), n, F, topn, q_begin, q_end, nbr.data_ptr(),
                          dist.data_ptr() if dist is not None else None, ws.data_ptr(), ws_bytes, stream)
    _lib.check(rc, "wsi_knn_topk")
    return (nbr, dist) if want_dist else nbr


def edge_pearson(feat: torch.Tensor, src: torch.Tensor, dst: torch.Tensor):
    """(sim fp32 [E], etype uint8 [E]) = Pearson r of the two feature rows of every edge; see wsi_edge_pearson."""
    lib = _lib.load()
    stream = _prep(feat)
    if feat.dtype != torch.float32 or feat.dim() != 2 or not feat.is_contiguous():
        raise TypeError("edge_pearson: features must be a contiguous fp32 [N, F] tensor")
    E = int(src.shape[0])
    sim = torch.empty(E, dtype=torch.float32, device=feat.device)
    et = torch.empty(E, dtype=torch.uint8, device=feat.device)
    rc = lib.wsi_edge_pearson(feat.data_ptr(), int(feat.shape[0]), int(feat.shape[1]), _vec(src, "src", torch.int64),
                              _vec(dst, "dst", torch.int64), E, sim.data_ptr(), et.data_ptr(), stream)
    _lib.check(rc, "wsi_edge_pearson")
    return sim, et
