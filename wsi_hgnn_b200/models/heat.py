"""HEATNet2 / HEATNet4 with the reference's constructor, forward() and state_dict contracts,
running on the sm_100a kernels of libwsi_hgnn.so.

Reference: models/HEATNet4.py:49-138 (HEATLayer), :141-247 (HEATNet4); models/HEATNet2.py:24-113
(same layer), :116-196 (HEATNet2).  Differences in HOW, not WHAT:
  * K/Q/V are projected once per node TYPE by one fused typed GEMM (the reference re-projects them
    for every relation, models/HEATNet4.py:95-102);
  * all relations of a layer run in ONE edge-attention launch over a relation-grouped CSR instead of a
    Python loop of DGL calls (models/HEATNet4.py:91-119);
  * the sigma(skip) mix, dropout mask and KeyError-passthrough are the epilogue of the a_linear GEMM
    (models/HEATNet4.py:122-136).
"""
import math
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..autograd import ALinearSkipFn, HeteroAttnFn, SegmentPoolFn, SkipMixFn, TypedLinearFn, _tc_chain
from ..hetero_graph import GraphPlan, HeteroGraph
from ._packing import PackCache, param_list, stack_linears

_POOLS = ("mean", "sum", "max")


def _check_pool(name: str) -> str:
    if name not in _POOLS:
        # 'att' (GlobalAttentionPooling) cannot take ntype= and raises in the reference too; anything else
        # is NotImplementedError there as well (models/HEATNet4.py:182-189).
        raise NotImplementedError(name)
    return name


def _graph_type_order(plan: GraphPlan, node_dict: Dict[str, int]):
    return [node_dict[nt] for nt in plan.ntypes]        # KeyError for an unknown type, as in the reference


def packed_features(G: HeteroGraph, plan: GraphPlan, h: Optional[Dict[str, torch.Tensor]], name: str = "feat"):
    """[N, F] type-major packed input features (G.nodes[nt].data['feat'] or the caller's dict)."""
    if h is None:
        # cached per plan, keyed by the identity + version of the per-type tensors: a later write to
        # G.nodes[nt].data['feat'] (feature masking, in-place normalisation) must be seen, as in the reference, which
        # re-reads the features on every forward (models/HEATNet4.py:198-206)
        sig = G.ndata_signature(name)
        hit = plan.cache.get(name)
        if hit is None or hit[0] != sig:
            hit = (sig, G.packed_ndata(name, torch.float32))
            plan.cache[name] = hit
        return hit[1]
    parts = [h[nt].to(torch.float32) for nt in plan.ntypes if G.num_nodes(nt) > 0]
    return torch.cat(parts, 0).contiguous()


def unpack_rows(plan: GraphPlan, x: torch.Tensor) -> Dict[str, torch.Tensor]:
    return {nt: x[plan.type_ptr[t]:plan.type_ptr[t + 1]] for t, nt in enumerate(plan.ntypes)}


def readout_scale(plan: GraphPlan, independent: bool) -> torch.Tensor:
    """[T*B] 0/1 mask of the readout rows: 0 where the reference emits a zero block instead of
    linears_prediction(pool) - a node type without nodes (models/HEATNet4.py:217-221,240); per graph
    for pack()ed graphs, per batch for dgl.batch semantics."""
    key = ("readout_scale", independent)
    if key not in plan.cache:
        if plan.device.type == "cuda":
            # from the device copy of the segment pointers: a pageable host -> device copy here is host-synchronous
            # and queues behind the next slide's 34 MB upload on the copy engine (it serialised the streamed path)
            T = len(plan.ntypes)
            ne = (plan.seg_ptr[1:] > plan.seg_ptr[:-1]).view(T, plan.B)
            m = ne if independent else ne.any(dim=1, keepdim=True).expand(T, plan.B)
            plan.cache[key] = m.to(torch.float32).reshape(-1).contiguous()
        else:
            ne = plan.seg_nonempty
            m = ne if independent else ne.any(dim=1, keepdim=True).expand_as(ne)
            plan.cache[key] = m.to(torch.float32).reshape(-1).contiguous().to(plan.device)
    return plan.cache[key]


class HEATLayer(nn.Module):
    """reference models/HEATNet4.py:49-138 (== models/HEATNet2.py:24-113)."""

    def __init__(self, in_size, out_size, node_dict, n_heads, dropout=0.2):
        super().__init__()
        self.weight = nn.Linear(in_size, out_size)     # "W_r": created, never used (models/HEATNet4.py:53-54)
        self.in_size = in_size
        self.out_size = out_size
        self.node_dict = node_dict
        self.num_ntypes = len(node_dict)
        self.n_heads = n_heads
        self.d_k = out_size // n_heads
        self.sqrt_dk = math.sqrt(self.d_k)
        T = self.num_ntypes
        self.k_linears = nn.ModuleList([nn.Linear(in_size, out_size) for _ in range(T)])
        self.q_linears = nn.ModuleList([nn.Linear(in_size, out_size) for _ in range(T)])
        self.v_linears = nn.ModuleList([nn.Linear(in_size, out_size) for _ in range(T)])
        self.a_linears = nn.ModuleList([nn.Linear(out_size, out_size) for _ in range(T)])
        self.e_linear = nn.Linear(1, 1)
        self.skip = nn.Parameter(torch.ones(T))
        self.drop = nn.Dropout(dropout)
        self._packs = PackCache()

    def _packed(self, order):
        D, H = self.out_size, self.n_heads
        perm = ops.head_perm(D, H)
        params = param_list(self, "all", self.parameters)

        def build():
            dev = self.skip.device
            pm = perm.to(dev) if perm is not None else None
            wk, bk = stack_linears(self.k_linears, order, row_perm=pm)
            wv, bv = stack_linears(self.v_linears, order, row_perm=pm)
            wq, bq = stack_linears(self.q_linears, order, row_perm=pm)
            w_kvq = torch.cat([wk, wv, wq], 1).contiguous()          # [T, 3D, in]
            b_kvq = torch.cat([bk, bv, bq], 1).contiguous()
            wa, ba = stack_linears(self.a_linears, order, col_perm=pm)
            skip = self.skip[torch.tensor(order, device=dev)].contiguous()
            return w_kvq, b_kvq, wa, ba, skip

        return self._packs.get(tuple(order), params, build) + (perm is not None,)

    def _packed_split(self, order):
        """Operand forms (ops.to_operand, the current matmul precision) of the K|V|Q and a_linear weight stacks."""
        w_kvq, b_kvq, wa, ba, skip, _ = self._packed(order)
        params = param_list(self, "all", self.parameters)
        opf = ops.matmul_opf()
        return self._packs.get(("split", opf, tuple(order)), params, lambda: (ops.to_operand(w_kvq, opf), ops.to_operand(wa, opf)))

    def tc_chain_ok(self, plan: GraphPlan) -> bool:
        """The pre-split tensor-core chain needs tile-friendly shapes and the lane-grouped attention layout."""
        D = self.out_size
        return (self.in_size == D and ops.head_perm(D, self.n_heads) is not None and
                ops.tc_ok(plan.N, D, 3 * D) and ops.tc_ok(plan.N, D, D))

    def forward_split(self, plan: GraphPlan, x: torch.Tensor, x_split: torch.Tensor, want_op: bool):
        """One layer on the pre-split chain: x fp32 [N, D] (residual) and its bf16 [hi; lo] form in, the same two
        out (x_split only if `want_op`, i.e. another layer follows).  No fp32 -> bf16 conversion pass: the
        attention kernel and the a_linear epilogue emit the split form directly."""
        D, H = self.out_size, self.n_heads
        order = _graph_type_order(plan, self.node_dict)
        w_kvq, b_kvq, wa, ba, skip, _ = self._packed(order)
        w_kvq_s, wa_s = self._packed_split(order)
        tpc = plan.type_ptr_c()
        kvq, _ = ops.typed_linear_op(x_split, w_kvq_s, b_kvq, plan.type_ptr, 3 * D, type_ptr_c=tpc)
        agg_s = ops.hetero_attn_work(kvq[:, :D], kvq[:, D:2 * D], kvq[:, 2 * D:], plan.attn_work(), plan.e_src,
                                     plan.e_sim, plan.e_rel, plan.node_inv_r, self.e_linear.weight,
                                     self.e_linear.bias, D, H, op_out=True)
        mask = None
        if self.training and self.drop.p > 0:
            mask = F.dropout(torch.ones((plan.N, D), dtype=torch.float32, device=x.device), self.drop.p, True)
        return ops.typed_linear_op(agg_s, wa_s, ba, plan.type_ptr, D, skip=skip, res=x, row_gate=plan.node_inv_r,
                                      drop_mask=mask, want_op=want_op, type_ptr_c=tpc)

    def forward_train(self, plan: GraphPlan, x: torch.Tensor) -> torch.Tensor:
        """Differentiable layer (autograd.py): the packed weights are built WITH grad tracking so that the gradients
        of the fused / permuted stacks flow back to the reference-shaped nn.Linear parameters."""
        D, H = self.out_size, self.n_heads
        perm = ops.head_perm(D, H)
        if perm is None:
            raise NotImplementedError(f"training needs the lane-grouped attention layout (D % 128 == 0, H a power of "
                                      f"two <= 32); got D={D}, H={H}")
        order = _graph_type_order(plan, self.node_dict)
        pm = perm.to(x.device)
        wk, bk = stack_linears(self.k_linears, order, row_perm=pm)
        wv, bv = stack_linears(self.v_linears, order, row_perm=pm)
        wq, bq = stack_linears(self.q_linears, order, row_perm=pm)
        wa, ba = stack_linears(self.a_linears, order, col_perm=pm)
        tpc = plan.type_ptr_c()
        kvq = TypedLinearFn.apply(x, torch.cat([wk, wv, wq], 1), torch.cat([bk, bv, bq], 1), plan.type_ptr, tpc)
        agg = HeteroAttnFn.apply(kvq, self.e_linear.weight, self.e_linear.bias, plan, D, H)        # HEATNet4.py:103-119
        skip_t = self.skip[torch.tensor(order, device=x.device)]
        if _tc_chain(plan.N, D, D):
            # a_linear + dropout + sigma(skip) mix + passthrough as ONE fused GEMM forward / one row kernel backward
            mask = None
            if self.training and self.drop.p > 0:
                mask = F.dropout(torch.ones((plan.N, D), dtype=torch.float32, device=x.device), self.drop.p, True)
            return ALinearSkipFn.apply(agg, wa, ba, x, skip_t, mask, plan.node_inv_r, plan.type_ptr, tpc)    # :134-135, :129-133
        lin = self.drop(TypedLinearFn.apply(agg, wa, ba, plan.type_ptr, tpc))                      # :134
        return SkipMixFn.apply(lin, x, skip_t, plan.type_ptr, plan.node_inv_r)                     # :135, passthrough :129-133

    def forward_packed(self, plan: GraphPlan, x: torch.Tensor) -> torch.Tensor:
        """x [N, in] type-major packed -> [N, out]."""
        D, H = self.out_size, self.n_heads
        order = _graph_type_order(plan, self.node_dict)
        w_kvq, b_kvq, wa, ba, skip, use_perm = self._packed(order)
        tpc = plan.type_ptr_c()
        kvq = ops.typed_linear(x, w_kvq, b_kvq, plan.type_ptr, type_ptr_c=tpc)
        if use_perm:                      # lane-grouped layout: hub-balanced work list
            agg = ops.hetero_attn_work(kvq[:, :D], kvq[:, D:2 * D], kvq[:, 2 * D:], plan.attn_work(), plan.e_src,
                                       plan.e_sim, plan.e_rel, plan.node_inv_r, self.e_linear.weight,
                                       self.e_linear.bias, D, H)
        else:
            plan.check()
            agg = ops.hetero_attn(kvq[:, :D], kvq[:, D:2 * D], kvq[:, 2 * D:], plan.rowptr, plan.e_src, plan.e_sim,
                                  plan.e_rel, plan.node_inv_r, self.e_linear.weight, self.e_linear.bias, D, H, False)
        mask = None
        if self.training and self.drop.p > 0:
            mask = F.dropout(torch.ones_like(agg), self.drop.p, True)
        return ops.typed_linear(agg, wa, ba, plan.type_ptr, skip=skip, res=x, row_gate=plan.node_inv_r,
                                drop_mask=mask, type_ptr_c=tpc)

    def forward(self, G: HeteroGraph, feat_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        plan = G.plan()
        x = packed_features(G, plan, feat_dict)
        return unpack_rows(plan, self.forward_packed(plan, x))


class _AttnParams(nn.Module):
    """Parameters of the reference's LinearAttentionBlock (models/HEATNet4.py:20-42).  The block is the
    identity on its first argument in every reachable call (softmax over an axis of length 1), so only
    the state_dict entry `attn.{k}.op.weight [1, 256, 1]` is kept."""

    def __init__(self):
        super().__init__()
        self.op = nn.Conv1d(256, 1, kernel_size=1, padding=0, bias=False)


class _HEATBase(nn.Module):
    def prepare_plan_begin(self, plan: GraphPlan):
        """The asynchronous half of prepare_plan (kernels + a non-blocking read of the totals): lets the caller enqueue
        other work before prepare_plan waits."""
        if len(self.gcs) and ops.head_perm(self.gcs[0].out_size, self.gcs[0].n_heads) is not None:
            plan.attn_work_begin()

    def prepare_plan(self, plan: GraphPlan):
        """Build ahead of time the plan-side structures the forward needs that cost a host read (the hub-balancing
        work list): the streaming evaluator calls this on its planning stream, one slide ahead."""
        if len(self.gcs) and ops.head_perm(self.gcs[0].out_size, self.gcs[0].n_heads) is not None:
            plan.attn_work()
        else:
            plan.check()

    # ------------------------------------------------------------------ native whole-forward driver
    def _native_ok(self, G: HeteroGraph, plan: GraphPlan, h, n_out: int) -> bool:
        """The one-call C driver (wsi_heat_forward) covers inference on the tensor-core chain with the fused readout."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in param_list(self, "all", self.parameters)):
            return False
        if h is not None or getattr(self, "explicit_heads", False) or n_out > ops.AFFINE_MAX_OUT or plan.N == 0:
            return False
        if any(l.training and l.drop.p > 0 for l in self.gcs):
            return False
        F_in, D = self.adapt_ws[0].in_features, self.adapt_ws[0].out_features
        return (len(self.gcs) > 0 and ops.tc_ok(plan.N, F_in, D) and all(l.tc_chain_ok(plan) for l in self.gcs))

    def _native_params(self, plan: GraphPlan, order, names, collapse_heads: bool):
        """struct wsi_heat_params for this (graph type order, weights version); rebuilt when a parameter changes."""
        import ctypes
        from .. import _lib
        params = param_list(self, "all", self.parameters)
        opf = ops.matmul_opf()

        def build():
            w_in, b_in = stack_linears(self.adapt_ws, order)
            keep = [ops.to_operand(w_in, opf), b_in]
            L = len(self.gcs)
            arr = lambda: (ctypes.c_void_p * max(L, 1))()
            w_kvq, b_kvq, w_a, b_a, skip, e_w, e_b = arr(), arr(), arr(), arr(), arr(), arr(), arr()
            for i, layer in enumerate(self.gcs):
                wk, bk, wa, ba, sk, _ = layer._packed(order)
                wks, was = layer._packed_split(order)
                ew, eb = layer.e_linear.weight.detach().reshape(-1).contiguous(), layer.e_linear.bias.detach().reshape(-1).contiguous()
                keep += [wks, bk, was, ba, sk, ew, eb]
                w_kvq[i], b_kvq[i], w_a[i], b_a[i], skip[i] = wks.data_ptr(), bk.data_ptr(), was.data_ptr(), ba.data_ptr(), sk.data_ptr()
                e_w[i], e_b[i] = ew.data_ptr(), eb.data_ptr()
            M, c, b_total = self._affine_maps(names, collapse_heads)
            keep += [M, c, b_total, w_kvq, b_kvq, w_a, b_a, skip, e_w, e_b]
            P = _lib.HeatParams()
            P.F, P.D, P.H, P.L = w_in.shape[2], w_in.shape[1], self.gcs[0].n_heads, L
            P.opf = opf
            P.w_in_split, P.b_in = keep[0].data_ptr(), b_in.data_ptr()
            P.w_kvq_split, P.b_kvq, P.w_a_split, P.b_a, P.skip, P.e_w, P.e_b = w_kvq, b_kvq, w_a, b_a, skip, e_w, e_b
            P.pool_op, P.n_out = ops.POOL_OPS[self.graph_pooling_type], M.shape[1]
            P.M, P.c = M.data_ptr(), (c.data_ptr() if c is not None else None)
            P.b_total = b_total.data_ptr() if b_total is not None else None
            return P, keep

        return self._packs.get(("native", opf, tuple(order), tuple(names), collapse_heads), params, build)

    def _forward_native(self, G: HeteroGraph, plan: GraphPlan, collapse_heads: bool, return_embeddings: bool):
        return self._forward_native_core(plan, packed_features(G, plan, None), G.independent, collapse_heads,
                                         return_embeddings)

    def forward_planned(self, plan: GraphPlan, feat: torch.Tensor, independent: bool = False):
        """Inference from an already built plan and the packed [N, F] features (no HeteroGraph): the entry the
        streaming evaluator uses.  -> logits, or None when the one-call driver does not take this shape / mode (the
        caller then goes through forward(G))."""
        n_out = self.head.out_features if hasattr(self, "head") else next(iter(self.linears_prediction.values())).out_features
        if not self.native_forward or not self._native_ok(None, plan, None, n_out):
            return None
        return self._forward_native_core(plan, feat, independent, hasattr(self, "head"), False)

    def stream_native(self, slides, dev, depth: int, ctx: Dict):
        """Logits of a LIST of flat slides through wsi_stream_forward (the whole pipelined loop in one C call).
        -> list of [1, out] pinned host tensors, or None when a slide is not the one-call driver's shape (the caller
        then uses the Python-issued pipeline)."""
        import ctypes
        from types import SimpleNamespace
        from .. import _lib
        lib = _lib.load()
        n = len(slides)
        if n == 0 or not self.native_forward:
            return None
        names0 = list(slides[0].header["ntypes"])
        n_out = self.head.out_features if hasattr(self, "head") else next(iter(self.linears_prediction.values())).out_features
        F_model = self.adapt_ws[0].in_features
        for s in slides:
            h = s.header
            N, E = sum(h["num_nodes"]), sum(h["num_edges"])
            if (N == 0 or E == 0 or list(h["ntypes"]) != names0 or h["feat_dim"] != F_model or h["feat_name"] != "feat"
                    or not s.blob.is_pinned() or not self._native_ok(None, SimpleNamespace(N=N), None, n_out)):
                return None
        order = [self.node_dict[nt] for nt in names0]
        P0, _keep = self._native_params(None, order, names0, hasattr(self, "head"))
        fp16_feat = [s.feat_is_fp16() for s in slides]
        if any(fp16_feat) and P0.opf != ops.OPF_F16:
            return None
        P = _lib.HeatParams.from_buffer_copy(P0)
        T = len(names0)
        tix = {nt: i for i, nt in enumerate(names0)}
        arr = (_lib.StreamSlide * n)()
        keep = []
        max_nb = max_n = max_e = max_r = 0
        for i, s in enumerate(slides):
            h = s.header
            N, E, R = sum(h["num_nodes"]), sum(h["num_edges"]), len(h["canonical_etypes"])
            ints = (ctypes.c_int32 * (T + 3 * R))(*h["num_nodes"], *h["num_edges"],
                                                 *[tix[ce[0]] for ce in h["canonical_etypes"]],
                                                 *[tix[ce[2]] for ce in h["canonical_etypes"]])
            keep.append(ints)
            base = ctypes.addressof(ints)
            a = arr[i]
            a.blob_host, a.nbytes = s.blob.data_ptr(), h["nbytes"]
            off = h["off"]
            a.off_feat, a.off_src, a.off_dst, a.off_sim = off["feat"], off["src"], off["dst"], off["sim"]
            a.n_nodes, a.n_edges, a.T, a.R, a.F, a.feat_is_op = N, E, T, R, F_model, 1 if fp16_feat[i] else 0
            a.nodes_per_type_host, a.edges_per_rel_host = base, base + 4 * T
            a.rel_src_type_host, a.rel_dst_type_host = base + 4 * (T + R), base + 4 * (T + 2 * R)
            max_nb, max_n, max_e, max_r = max(max_nb, h["nbytes"]), max(max_n, N), max(max_e, E), max(max_r, R)
        depth = max(3, min(int(depth), 16))
        slot = lib.wsi_stream_slot_bytes(max_nb, max_n, max_e, F_model, P.D, T, max_r, P.n_out)
        hslot = lib.wsi_stream_host_slot_bytes(max_n, T, max_r)
        ws = ctx.get("native_ws")
        if ws is None or ws[0].numel() < depth * slot + 256 or ws[1].numel() < depth * hslot:
            ws = (torch.empty(int(depth * slot * 1.25) + 256, dtype=torch.uint8, device=dev),
                  torch.empty(int(depth * hslot * 1.25), dtype=torch.uint8).pin_memory())
            ctx["native_ws"] = ws
        logits = torch.empty((n, P.n_out), dtype=torch.float32).pin_memory()
        stream = ops._prep(ws[0])
        rc = lib.wsi_stream_forward(arr, n, ctypes.byref(P), logits.data_ptr(), depth, ws[0].data_ptr(), ws[0].numel(),
                                    ws[1].data_ptr(), ws[1].numel(), stream)
        _lib.check(rc, "wsi_stream_forward")
        return [logits[i:i + 1] for i in range(n)]

    def slide_plan_native(self, slide, blob: torch.Tensor, head: torch.Tensor, slot: Dict, plan_stream):
        """Phase 1 of blob -> logits of ONE flat slide (wsi_slide_plan: CSR + work-list counting enqueued on `plan_stream`,
        totals on their way to pinned host memory; nothing waits).  `head` = the slide's plan head (FlatSlide._plan_head)
        on the device, `slot` = per-buffer scratch ({'ws', 'totals'}) recycled by the caller once the slide's forward has
        finished.  -> an opaque state for slide_run_native, or None when the shapes are not the driver's (the caller then
        takes the generic path)."""
        import ctypes
        from types import SimpleNamespace
        from .. import _lib
        lib = _lib.load()
        hd, hdr = slide._plan_head(), slide.header
        N, E, T, R, F = hd["N"], hd["E"], hd["T"], hd["R"], hdr["feat_dim"]
        n_out = self.head.out_features if hasattr(self, "head") else next(iter(self.linears_prediction.values())).out_features
        shape = SimpleNamespace(N=N)
        if (not self.native_forward or N == 0 or E == 0 or not self._native_ok(None, shape, None, n_out)
                or F != self.adapt_ws[0].in_features or hdr["feat_name"] != "feat"):
            return None
        names = list(hd["ntypes"])
        order = [self.node_dict[nt] for nt in names]
        P0, _keep = self._native_params(None, order, names, hasattr(self, "head"))
        if slide.feat_is_fp16() and P0.opf != ops.OPF_F16:
            return None                                         # fp16 features + another operand format: generic path
        P = _lib.HeatParams.from_buffer_copy(P0)                # per-slide copy: seg_scale points into this slide's head
        D = P.D
        key = (N, E, F, D, T)
        if slot.get("key") != key:
            ws_bytes = lib.wsi_slide_forward_workspace_bytes(N, E, F, D, T, E)
            slot["ws"] = torch.empty(ws_bytes, dtype=torch.uint8, device=blob.device)
            slot["ws_bytes"], slot["key"] = ws_bytes, key
            slot["totals"] = torch.zeros(4, dtype=torch.int32).pin_memory()
        off, base = hdr["off"], blob.data_ptr()
        hp = head.data_ptr()
        d = _lib.SlideDesc()
        d.feat, d.ldf = base + off["feat"], F
        d.feat_is_op = 1 if slide.feat_is_fp16() else 0
        d.src, d.dst, d.sim = base + off["src"], base + off["dst"], base + off["sim"]
        d.seg_ptr = hp
        d.rel_table = hp + 4 * hd["n0"]
        d.node_inv_r = hp + 4 * hd["n1p"]
        tpc = ops.host_i32(hd["type_ptr"])
        d.type_ptr_host = ctypes.cast(tpc, ctypes.c_void_p)
        d.n_nodes, d.n_edges, d.T, d.R, d.chunk = N, E, T, R, 16
        P.seg_scale = hp + 4 * hd["n1"]
        ops._prep(blob)
        rc = lib.wsi_slide_plan(ctypes.byref(d), ctypes.byref(P), E, slot["totals"].data_ptr(), slot["ws"].data_ptr(),
                                slot["ws_bytes"], plan_stream.cuda_stream)
        _lib.check(rc, "wsi_slide_plan")
        return (d, P, tpc, _keep, E, blob, head)

    def slide_run_native(self, state, slot: Dict, plan_stream, main_stream) -> torch.Tensor:
        """Phase 2 (wsi_slide_run): the caller has waited for the event it recorded on `plan_stream` after phase 1, so the
        totals are on the host; work-list fill + the whole forward are enqueued.  -> logits [1, out] (on `main_stream`)."""
        import ctypes
        from .. import _lib
        lib = _lib.load()
        d, P, _tpc, _keep, E, blob, _head = state
        ops._prep(blob)
        logits = torch.empty((1, P.n_out), dtype=torch.float32, device=blob.device)
        rc = lib.wsi_slide_run(ctypes.byref(d), ctypes.byref(P), E, slot["totals"].data_ptr(), logits.data_ptr(), P.n_out,
                               slot["ws"].data_ptr(), slot["ws_bytes"], plan_stream.cuda_stream, main_stream.cuda_stream)
        _lib.check(rc, "wsi_slide_run")
        return logits

    def slide_forward_native(self, slide, blob: torch.Tensor, head: torch.Tensor, slot: Dict, plan_stream, main_stream):
        """Both phases back to back with one host wait on `plan_stream` in between (un-pipelined callers)."""
        state = self.slide_plan_native(slide, blob, head, slot, plan_stream)
        if state is None:
            return None
        plan_stream.synchronize()
        return self.slide_run_native(state, slot, plan_stream, main_stream)

    def _forward_native_core(self, plan: GraphPlan, feat: torch.Tensor, independent: bool, collapse_heads: bool,
                             return_embeddings: bool):
        import ctypes
        from .. import _lib
        lib = _lib.load()
        order = _graph_type_order(plan, self.node_dict)
        names = list(plan.ntypes)
        P, _keep = self._native_params(plan, order, names, collapse_heads)
        key = ("native_graph", independent)
        if key not in plan.cache:
            w = plan.attn_work()
            g = _lib.HeatGraph()
            g.n_rows, g.T, g.B = plan.N, len(names), plan.B
            tpc = plan.type_ptr_c()
            g.type_ptr_host = ctypes.cast(tpc, ctypes.c_void_p)
            g.seg_ptr = plan.seg_ptr.data_ptr()
            g.e_src, g.e_sim, g.e_rel, g.node_inv_r = (plan.e_src.data_ptr(), plan.e_sim.data_ptr(), plan.e_rel.data_ptr(),
                                                       plan.node_inv_r.data_ptr())
            g.items, g.n_items = w["items"].data_ptr(), w["n_items"]
            g.split_row, g.split_ptr, g.part_rel = w["split_row"].data_ptr(), w["split_ptr"].data_ptr(), w["part_rel"].data_ptr()
            g.part_split = w["part_split"].data_ptr() if w.get("part_split") is not None else None
            g.split_cnt = w["split_cnt"].data_ptr() if w.get("split_cnt") is not None else None
            g.sched = w["sched"].data_ptr() if w.get("sched") is not None else None
            g.n_split, g.n_part = w["n_split"], w["n_part"]
            scale = readout_scale(plan, independent)
            ws_bytes = lib.wsi_heat_forward_workspace_bytes(plan.N, P.F, P.D, w["n_part"], len(names), plan.B)
            plan.cache[key] = (g, scale, ws_bytes, tpc, w)
        g, scale, ws_bytes, _, _ = plan.cache[key]
        feat_is_op = 0
        if feat.dtype == torch.float16 and P.opf == ops.OPF_F16 and feat.is_contiguous() and feat.data_ptr() % 128 == 0:
            feat_is_op = 1                                      # fp16 features of a flat slide: already the GEMM's operand
        elif feat.dtype != torch.float32 or feat.stride(1) != 1:
            feat = feat.float().contiguous()
        stream = ops._prep(feat)
        P.seg_scale = scale.data_ptr()
        dev = feat.device
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        logits = torch.empty((plan.B, P.n_out), dtype=torch.float32, device=dev)
        x_out = torch.empty((plan.N, P.D), dtype=torch.float32, device=dev) if return_embeddings else None
        rc = lib.wsi_heat_forward(feat.data_ptr(), feat.stride(0), feat_is_op, ctypes.byref(g), ctypes.byref(P),
                                  x_out.data_ptr() if x_out is not None else None, P.D, logits.data_ptr(), P.n_out,
                                  ws.data_ptr(), ws_bytes, stream)
        _lib.check(rc, "wsi_heat_forward")
        return (logits, unpack_rows(plan, x_out)) if return_embeddings else logits

    def _trunk(self, G: HeteroGraph, h):
        plan = G.plan()
        order = _graph_type_order(plan, self.node_dict)
        x = packed_features(G, plan, h)
        params = param_list(self, "in", lambda: (p for m in self.adapt_ws for p in m.parameters()))
        w_in, b_in = self._packs.get(("in", tuple(order)), params, lambda: stack_linears(self.adapt_ws, order))
        if torch.is_grad_enabled() and any(p.requires_grad for p in param_list(self, "all", self.parameters)):
            w_in_g, b_in_g = stack_linears(self.adapt_ws, order)                             # with grad tracking
            x = TypedLinearFn.apply(x, w_in_g, b_in_g, plan.type_ptr, plan.type_ptr_c())
            for layer in self.gcs:
                x = layer.forward_train(plan, x)
            return plan, x
        F_in, D = int(x.shape[1]), int(w_in.shape[1])
        if len(self.gcs) > 0 and ops.tc_ok(plan.N, F_in, D) and all(l.tc_chain_ok(plan) for l in self.gcs):
            # tensor-core chain on operands kept in operand form (ops.to_operand): one conversion pass for the raw
            # features, every later operand is emitted in that form by the kernel that produces it
            opf = ops.matmul_opf()
            w_in_s = self._packs.get(("in_split", opf, tuple(order)), params, lambda: ops.to_operand(w_in, opf))
            x, xs = ops.typed_linear_op(ops.to_operand(x), w_in_s, b_in, plan.type_ptr, D, want_op=True,
                                           type_ptr_c=plan.type_ptr_c())                     # HEATNet4.py:198-206
            for i, layer in enumerate(self.gcs):                                             # :213-214
                x, xs = layer.forward_split(plan, x, xs, want_op=i + 1 < len(self.gcs))
            return plan, x
        x = ops.typed_linear(x, w_in, b_in, plan.type_ptr, type_ptr_c=plan.type_ptr_c())    # HEATNet4.py:198-206
        for layer in self.gcs:                                                               # :213-214
            x = layer.forward_packed(plan, x)
        return plan, x

    def _readout(self, G, plan, x):
        """[T*B, n_pred] = linears_prediction[type](pool_type(x)) with the empty-type zero block."""
        pooled = ops.segment_pool(x, plan.seg_ptr, len(plan.ntypes) * plan.B, self.graph_pooling_type)
        return self._predict_pooled(G, plan, pooled)

    def _predict_pooled(self, G, plan, pooled):
        """[T*B, D] pooled (type-major) -> [T*B, n_pred] = linears_prediction[type](pooled) with the empty-type zero block."""
        names = list(plan.ntypes)
        params = param_list(self, ("pred", tuple(names)), lambda: (p for nt in names for p in self.linears_prediction[nt].parameters()))

        def build():
            ws = torch.stack([self.linears_prediction[nt].weight for nt in names]).contiguous()
            bs = torch.stack([self.linears_prediction[nt].bias for nt in names]).contiguous()
            return ws, bs

        w_p, b_p = self._packs.get(("pred", tuple(names)), params, build)
        return ops.typed_linear(pooled, w_p, b_p, plan.readout_ptr(), row_scale=readout_scale(plan, G.independent))


    def _readout_train(self, G, plan, x):
        """[T*B, n_pred] = linears_prediction[type](pool_type(x)) with the empty-type zero block, differentiable."""
        names = list(plan.ntypes)
        pooled = SegmentPoolFn.apply(x, plan, len(names) * plan.B, self.graph_pooling_type)
        w_p = torch.stack([self.linears_prediction[nt].weight for nt in names])
        b_p = torch.stack([self.linears_prediction[nt].bias for nt in names])
        o = TypedLinearFn.apply(pooled, w_p, b_p, plan.readout_ptr(), None)
        return o * readout_scale(plan, G.independent).unsqueeze(1)

    def _affine_maps(self, names, collapse_heads: bool):
        """(M [T, out, D], c [T, out], b_total [out] | None): the per-type affine maps of the readout (see _readout_affine)."""
        T = len(names)
        heads = [self.head_2, self.head_1, self.head] if collapse_heads else []
        params = param_list(self, ("affine", tuple(names)), lambda: (
            [p for nt in names for p in self.linears_prediction[nt].parameters()] + [p for h in heads for p in h.parameters()]))

        def build():
            wp = torch.stack([self.linears_prediction[nt].weight for nt in names]).double()      # [T, n_pred, D]
            bp = torch.stack([self.linears_prediction[nt].bias for nt in names]).double()        # [T, n_pred]
            if not collapse_heads:
                return wp.float().contiguous(), bp.float().contiguous(), None
            w2, b2 = self.head_2.weight.double(), self.head_2.bias.double()
            w1, b1 = self.head_1.weight.double(), self.head_1.bias.double()
            wh, bh = self.head.weight.double(), self.head.bias.double()
            wc = wh @ w1 @ w2                                                                    # [out, 256 T]
            b_total = wh @ (w1 @ b2 + b1) + bh
            blocks = wc.view(wc.shape[0], T, -1).permute(1, 0, 2)                                # [T, out, 256]
            M = torch.bmm(blocks, wp)                                                            # [T, out, D]
            c = torch.bmm(blocks, bp.unsqueeze(-1)).squeeze(-1)                                  # [T, out]
            return M.float().contiguous(), c.float().contiguous(), b_total.float().contiguous()

        return self._packs.get(("affine", tuple(names)), params, build)

    def _readout_affine(self, G, plan, x, collapse_heads: bool):
        """[B, out_dim] logits by the fused pool + affine kernel.  HEATNet2: M_t = linears_prediction[t]
        (models/HEATNet2.py:181-194).  HEATNet4: the chain linears_prediction -> cat -> head_2 -> head_1 -> head
        (models/HEATNet4.py:216-245) contains no nonlinearity (and LinearAttentionBlock is the identity), so it equals
        one [out, D] map per node type plus a constant; the composite is formed in fp64 on the host side of the pack
        cache and rebuilt whenever one of the parameters changes."""
        names = list(plan.ntypes)
        T, B = len(names), plan.B
        M, c, b_total = self._affine_maps(names, collapse_heads)
        return ops.segment_pool_affine(x, plan.seg_ptr, T, B, self.graph_pooling_type, M, c, b_total,
                                       readout_scale(plan, G.independent))


class HEATNet4(_HEATBase):
    """reference models/HEATNet4.py:141-247.  forward(G, h=None) -> logits [B, out_dim]."""

    def __init__(self, in_dim, hidden_dim, out_dim, n_layers, n_heads, node_dict, dropuout,
                 graph_pooling_type="mean"):
        super().__init__()
        self.node_dict = node_dict
        self.n_layers = n_layers
        self.graph_pooling_type = _check_pool(graph_pooling_type)
        self.linears_prediction = nn.ModuleDict({k: nn.Linear(hidden_dim, 256) for k in node_dict})
        self.adapt_ws = nn.ModuleList([nn.Linear(in_dim, hidden_dim) for _ in node_dict])
        self.gcs = nn.ModuleList([HEATLayer(hidden_dim, hidden_dim, node_dict, n_heads, dropuout)
                                  for _ in range(n_layers)])
        self.attn = nn.ModuleDict({k: _AttnParams() for k in node_dict})
        self.head_2 = nn.Linear(256 * len(node_dict), 256)
        self.head_1 = nn.Linear(256, 64)
        self.head = nn.Linear(64, out_dim)
        self.explicit_heads = False      # True: run linears_prediction / head_2 / head_1 / head as separate GEMMs
        self.native_forward = True       # inference through the one-call C driver (wsi_heat_forward) when it applies
        self._packs = PackCache()

    def forward(self, G: HeteroGraph, h=None, return_embeddings: bool = False):
        if self.native_forward and G.device.type == "cuda":
            plan = G.plan()
            if self._native_ok(G, plan, h, self.head.out_features):     # inference: the whole chain in one host call
                return self._forward_native(G, plan, True, return_embeddings)
        plan, x = self._trunk(G, h)
        T, B = len(plan.ntypes), plan.B
        if x.requires_grad:                                             # training: explicit, differentiable chain
            o = self._readout_train(G, plan, x)                         # [T*B, 256]      :216-240
            z = o.view(T, B, 256).permute(1, 0, 2).reshape(B, T * 256)  # cat(dim=1) in G.ntypes order
            one = [0, B]
            for head in (self.head_2, self.head_1, self.head):          # :243-245
                z = TypedLinearFn.apply(z, head.weight.unsqueeze(0), head.bias.unsqueeze(0), one, None)
            return (z, unpack_rows(plan, x)) if return_embeddings else z
        if self.head.out_features <= ops.AFFINE_MAX_OUT and not self.explicit_heads:
            g = self._readout_affine(G, plan, x, collapse_heads=True)   # :216-245 as one fused launch pair
            return (g, unpack_rows(plan, x)) if return_embeddings else g
        g = self._heads(self._readout(G, plan, x), T, B)                # :216-245
        return (g, unpack_rows(plan, x)) if return_embeddings else g

    def _heads(self, o, T, B):
        """[T*B, 256] per-type predictions -> cat(dim=1) in G.ntypes order -> head_2 -> head_1 -> head (:240-245)."""
        z = o if B == 1 else o.view(T, B, 256).permute(1, 0, 2).contiguous()
        z = z.view(B, T * 256)
        one = [0, B]
        z = ops.typed_linear(z, self.head_2.weight.unsqueeze(0), self.head_2.bias.unsqueeze(0), one)   # :243
        z = ops.typed_linear(z, self.head_1.weight.unsqueeze(0), self.head_1.bias.unsqueeze(0), one)   # :244
        return ops.typed_linear(z, self.head.weight.unsqueeze(0), self.head.bias.unsqueeze(0), one)    # :245

    def logits_from_pooled(self, G, plan, pooled):
        """[T*B, D] typed readout (already reduced, e.g. over the ranks of a node-sharded slide) -> logits [B, out]."""
        return self._heads(self._predict_pooled(G, plan, pooled), len(plan.ntypes), plan.B)


class HEATNet2(_HEATBase):
    """reference models/HEATNet2.py:116-196.  forward(G, h=None) -> logits [B, out_dim]."""

    def __init__(self, in_dim, hidden_dim, out_dim, n_layers, n_heads, node_dict, dropuout,
                 graph_pooling_type="mean"):
        super().__init__()
        self.node_dict = node_dict
        self.n_layers = n_layers
        self.graph_pooling_type = _check_pool(graph_pooling_type)
        self.linears_prediction = nn.ModuleDict({k: nn.Linear(hidden_dim, out_dim) for k in node_dict})
        self.adapt_ws = nn.ModuleList([nn.Linear(in_dim, hidden_dim) for _ in node_dict])
        self.gcs = nn.ModuleList([HEATLayer(hidden_dim, hidden_dim, node_dict, n_heads, dropuout)
                                  for _ in range(n_layers)])
        self.native_forward = True       # inference through the one-call C driver (wsi_heat_forward) when it applies
        self._packs = PackCache()

    def forward(self, G: HeteroGraph, h=None, return_embeddings: bool = False):
        if self.native_forward and G.device.type == "cuda":
            plan = G.plan()
            n_pred = next(iter(self.linears_prediction.values())).out_features
            if self._native_ok(G, plan, h, n_pred):                     # inference: the whole chain in one host call
                return self._forward_native(G, plan, False, return_embeddings)
        plan, x = self._trunk(G, h)
        T, B = len(plan.ntypes), plan.B
        if x.requires_grad:                                             # training: differentiable readout
            g = self._readout_train(G, plan, x).view(T, B, -1).sum(0)   # HEATNet2.py:181-194
            return (g, unpack_rows(plan, x)) if return_embeddings else g
        n_pred = next(iter(self.linears_prediction.values())).out_features
        if n_pred <= ops.AFFINE_MAX_OUT:
            g = self._readout_affine(G, plan, x, collapse_heads=False)  # HEATNet2.py:181-194
            return (g, unpack_rows(plan, x)) if return_embeddings else g
        o = self._readout(G, plan, x)                                   # [T*B, out]   HEATNet2.py:181-194
        g = o.view(T, B, -1).sum(0)
        return (g, unpack_rows(plan, x)) if return_embeddings else g

    def logits_from_pooled(self, G, plan, pooled):
        """[T*B, D] typed readout (already reduced over the ranks of a node-sharded slide) -> logits [B, out]."""
        return self._predict_pooled(G, plan, pooled).view(len(plan.ntypes), plan.B, -1).sum(0)
