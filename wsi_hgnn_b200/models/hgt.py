"""HGT with the reference's constructor, forward() and state_dict contracts on the sm_100a kernels.

Reference: models/HGT.py:21-127 (HGTLayer), :130-209 (HGT).  Same maths, different schedule:
  * K/Q/V once per node type (fused typed GEMM) instead of once per relation (models/HGT.py:75-84);
  * the per-relation d_k x d_k maps relation_att / relation_msg (models/HGT.py:88-93) are applied to
    (dst, relation) SEGMENTS - the query side  <Q, K A> = <A Q, K>  before the edge kernel and the message
    side  sum_e a_e (V M) = (sum_e a_e V) M  after it - instead of to every source node x relation;
  * one edge-attention launch over all segments (models/HGT.py:95-106), then the cross-relation mean.
"""
import math
from typing import Dict

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..autograd import (RelTransformFn, SegAttnFn, SegmentCombineFn, SegmentPoolFn, SkipMixFn, TypedLinearFn)
from ..hetero_graph import GraphPlan, HeteroGraph
from ._packing import PackCache, _Permute, param_list, stack_linears
from .heat import _check_pool, _graph_type_order, packed_features, readout_scale, unpack_rows


def _relation_groups(plan: GraphPlan, edge_dict: Dict, tag):
    """Per (graph, model) tensors that order the (dst, relation) segments by MODEL relation id."""
    key = ("hgt_groups", tag)
    if key not in plan.cache:
        segs = plan.segments()
        R = len(edge_dict)
        slot2model = torch.tensor([edge_dict[ce] for ce in plan.rel_list] or [0], dtype=torch.int64,
                                  device=plan.device)          # KeyError for an unknown relation (models/HGT.py:86)
        seg_rel = slot2model[segs["seg_slot"]] if segs["S"] else torch.zeros(0, dtype=torch.int64, device=plan.device)
        order = torch.argsort(seg_rel, stable=True)
        counts = torch.bincount(seg_rel, minlength=R)
        rel_ptr = [0] + torch.cumsum(counts, 0).tolist()
        # relation-sorted view of the segments (the tensor-core relation transforms work on rows grouped by relation):
        # work item i = segment order[i] -> (row i, its edge range); seg_pos = where segment s sits in that order
        sp = segs["seg_ptr"].to(torch.int64)
        S = int(segs["S"])
        items = torch.stack([torch.arange(S, device=plan.device), sp[:-1][order], sp[1:][order],
                             torch.full((S,), -1, dtype=torch.int64, device=plan.device)], 1).to(torch.int32).contiguous() \
            if S else torch.zeros((0, 4), dtype=torch.int32, device=plan.device)
        seg_pos = torch.empty(S, dtype=torch.int64, device=plan.device)
        seg_pos[order] = torch.arange(S, device=plan.device)
        # the SEGMENT GRAPH in that order (rows = segments, one relation run per row): the hub-balancing work list and
        # the kernels of HEAT's edge attention run on it unchanged (k-NN in-degrees are heavy-tailed: one warp per
        # segment leaves the launch waiting for its largest hub)
        lens = (sp[1:] - sp[:-1])[order]
        sg_ptr = torch.zeros(S + 1, dtype=torch.int64, device=plan.device)
        sg_ptr[1:] = torch.cumsum(lens, 0)
        e_perm = torch.repeat_interleave(sp[:-1][order] - sg_ptr[:-1], lens) + torch.arange(plan.E, device=plan.device)
        seg_graph = None
        if S and plan.device.type == "cuda":
            sg_rowptr = sg_ptr.to(torch.int32).contiguous()
            sg_rel = torch.zeros(plan.E, dtype=torch.uint8, device=plan.device)
            seg_graph = dict(rowptr=sg_rowptr, e_src=plan.e_src[e_perm].contiguous(), e_rel=sg_rel,
                             e_sim=torch.ones(plan.E, dtype=torch.float32, device=plan.device),
                             inv=torch.ones(S, dtype=torch.float32, device=plan.device),
                             ew=torch.ones(1, dtype=torch.float32, device=plan.device),
                             eb=torch.zeros(1, dtype=torch.float32, device=plan.device),
                             work=ops.plan_attn_work_finish(ops.plan_attn_work_begin(sg_rowptr, sg_rel, S, 16)))
        plan.cache[key] = dict(
            seg_graph=seg_graph,
            seg_rel=seg_rel.to(torch.int32).contiguous(),
            order=order.to(torch.int32).contiguous(),
            dst_of_order=segs["seg_dst"][order].contiguous(),
            items=items, seg_rel_sorted=seg_rel[order].to(torch.int32).contiguous(),
            seg_pos=seg_pos.to(torch.int32).contiguous(),
            rel_ptr_c=ops.host_i32(rel_ptr), rel_ptr=rel_ptr, R=R)
    return plan.cache[key]


class HGTLayer(nn.Module):
    """reference models/HGT.py:21-127."""

    def __init__(self, in_dim, out_dim, node_dict, edge_dict, n_heads, dropout=0.2, use_norm=False):
        super().__init__()
        self.in_dim = in_dim
        self.out_dim = out_dim
        self.node_dict = node_dict
        self.edge_dict = edge_dict
        self.num_types = len(node_dict)
        self.num_relations = len(edge_dict)
        self.total_rel = self.num_types * self.num_relations * self.num_types
        self.n_heads = n_heads
        self.d_k = out_dim // n_heads
        self.sqrt_dk = math.sqrt(self.d_k)
        self.att = None
        self.use_norm = use_norm
        T = self.num_types
        self.k_linears = nn.ModuleList([nn.Linear(in_dim, out_dim) for _ in range(T)])
        self.q_linears = nn.ModuleList([nn.Linear(in_dim, out_dim) for _ in range(T)])
        self.v_linears = nn.ModuleList([nn.Linear(in_dim, out_dim) for _ in range(T)])
        self.a_linears = nn.ModuleList([nn.Linear(out_dim, out_dim) for _ in range(T)])
        self.norms = nn.ModuleList([nn.LayerNorm(out_dim) for _ in range(T)] if use_norm else [])
        self.relation_pri = nn.Parameter(torch.ones(self.num_relations, self.n_heads))
        self.relation_att = nn.Parameter(torch.Tensor(self.num_relations, n_heads, self.d_k, self.d_k))
        self.relation_msg = nn.Parameter(torch.Tensor(self.num_relations, n_heads, self.d_k, self.d_k))
        self.skip = nn.Parameter(torch.ones(T))
        self.drop = nn.Dropout(dropout)
        nn.init.xavier_uniform_(self.relation_att)
        nn.init.xavier_uniform_(self.relation_msg)
        self._packs = PackCache()

    def _packed(self, order):
        params = param_list(self, "all", self.parameters)

        def build():
            dev = self.skip.device
            wk, bk = stack_linears(self.k_linears, order)
            wv, bv = stack_linears(self.v_linears, order)
            wq, bq = stack_linears(self.q_linears, order)
            w_kvq = torch.cat([wk, wv, wq], 1).contiguous()
            b_kvq = torch.cat([bk, bv, bq], 1).contiguous()
            wa, ba = stack_linears(self.a_linears, order)
            skip = self.skip[torch.tensor(order, device=dev)].contiguous()
            if self.use_norm:
                gamma = torch.stack([self.norms[i].weight for i in order]).contiguous()
                beta = torch.stack([self.norms[i].bias for i in order]).contiguous()
            else:
                gamma = beta = None
            return w_kvq, b_kvq, wa, ba, skip, gamma, beta

        return self._packs.get(tuple(order), params, build)

    def forward_packed(self, plan: GraphPlan, x: torch.Tensor, x_op=None, want_op: bool = False):
        """-> h' [N, D]; with want_op -> (h', operand-form copy of h' or None): the tensor-core schedule hands the next
        layer its GEMM operand (x_op) so that no conversion pass runs between the layers."""
        out, out_op = self.forward_packed_pair(plan, x, x_op, want_op)
        return (out, out_op) if want_op else out

    def forward_packed_pair(self, plan: GraphPlan, x: torch.Tensor, x_op=None, want_op: bool = False):
        out = self._forward_packed(plan, x, x_op, want_op)
        return out if isinstance(out, tuple) else (out, None)

    def _forward_packed(self, plan: GraphPlan, x: torch.Tensor, x_op, want_op: bool):
        D, H, dk = self.out_dim, self.n_heads, self.d_k
        order = _graph_type_order(plan, self.node_dict)
        w_kvq, b_kvq, wa, ba, skip, gamma, beta = self._packed(order)
        tpc = plan.type_ptr_c()
        segs = plan.segments()
        grp = _relation_groups(plan, self.edge_dict, id(self.edge_dict))
        S = segs["S"]
        opf = ops.matmul_opf()
        if (ops.head_perm(D, H) is not None and ops.tc_ok(S, D, D) and ops.tc_ok(plan.N, self.in_dim, 2 * D)
                and ops.tc_ok(plan.N, self.in_dim, D) and ops.tc_ok(plan.N, D, D)):
            return self._forward_packed_tc(plan, x, order, grp, segs, opf, x_op, want_op)
        if opf == ops.OPF_BF16 and ops.tc_ok(plan.N, self.in_dim, 2 * D) and ops.tc_ok(plan.N, self.in_dim, D):
            # bf16 storage of K | V (set_matmul_precision("bf16"): BASELINE config 3): the K|V GEMM writes ONLY the bf16
            # operand-form copy, the edge kernel gathers half the bytes; Q stays fp32 for the relation transform.
            # Accumulation is fp32 everywhere.
            params = param_list(self, "all", self.parameters)
            w_kv_op, w_q_op = self._packs.get(("kv16", opf, tuple(order)), params, lambda: (
                ops.to_operand(w_kvq[:, :2 * D].contiguous(), opf), ops.to_operand(w_kvq[:, 2 * D:].contiguous(), opf)))
            xs = ops.to_operand(x, opf)
            _, kv = ops.typed_linear_op(xs, w_kv_op, b_kvq[:, :2 * D].contiguous(), plan.type_ptr, 2 * D, want_y=False,
                                        want_op=True, type_ptr_c=tpc, opf=opf)
            q, _ = ops.typed_linear_op(xs, w_q_op, b_kvq[:, 2 * D:].contiguous(), plan.type_ptr, D, type_ptr_c=tpc, opf=opf)
            k, v = kv[:, :D], kv[:, D:]
        else:
            kvq = ops.typed_linear(x, w_kvq, b_kvq, plan.type_ptr, type_ptr_c=tpc)
            k, v, q = kvq[:, :D], kvq[:, D:2 * D], kvq[:, 2 * D:]
        # q'_seg = relation_att[r,h] . q[dst,h]   (w_kn=0: y_n = sum_k W[n,k] x_k)
        qseg = ops.rel_transform(q, grp["dst_of_order"], grp["order"], self.relation_att, grp["rel_ptr_c"], grp["R"],
                                 H, dk, False, S)
        aggseg = ops.hetero_attn_seg(k, v, qseg, segs["seg_ptr"], grp["seg_rel"], plan.e_src, self.relation_pri,
                                     D, H)
        # msg_seg = aggseg[h] . relation_msg[r,h]   (w_kn=1: y_n = sum_k x_k W[k,n])
        msgseg = ops.rel_transform(aggseg, grp["order"], grp["order"], self.relation_msg, grp["rel_ptr_c"], grp["R"],
                                   H, dk, True, S)
        agg = ops.segment_combine(msgseg, segs["row_seg_ptr"], plan.node_inv_r, plan.N, D)
        mask = None
        if self.training and self.drop.p > 0:
            mask = F.dropout(torch.ones_like(agg), self.drop.p, True)
        out = ops.typed_linear(agg, wa, ba, plan.type_ptr, skip=skip, res=x, row_gate=plan.node_inv_r,
                               drop_mask=mask, type_ptr_c=tpc)
        if self.use_norm:
            # NB the reference normalises only types that received a message (models/HGT.py:118-126);
            # passthrough types keep h unchanged, so the LayerNorm is applied row-gated below.
            out = _gated_layernorm(out, gamma, beta, plan, tpc)
        return out

    def _packed_tc(self, order, opf, dev):
        """Operand-form weights of the tensor-core schedule (cached per type order / operand format / parameter version):
        K | V projected straight into the lane-grouped column order of the edge kernel (row-permuted weights), the per
        (relation, head) d_k x d_k maps as ONE block-diagonal [D, D] matrix per relation with the column orders folded in:
          q'_phys = W_att_big[r] q        W_att_big[r][p, :]  = blockdiag_h(relation_att[r, h])[perm[p], :]
          msg     = W_msg_big[r] agg_phys  W_msg_big[r][:, p] = blockdiag_h(relation_msg[r, h]^T)[:, perm[p]]
        (3/4 of the block-diagonal product is zeros - tensor-core time well spent: one plain grouped GEMM per transform,
        no per-head launches, and the permutations cost nothing)."""
        params = param_list(self, "all", self.parameters)

        def build():
            D, H = self.out_dim, self.n_heads
            w_kvq, b_kvq, wa, ba, skip, gamma, beta = self._packed(order)
            pm = ops.head_perm(D, H).to(dev)
            w_kv = torch.cat([w_kvq[:, :D][:, pm], w_kvq[:, D:2 * D][:, pm]], 1).contiguous()
            b_kv = torch.cat([b_kvq[:, :D][:, pm], b_kvq[:, D:2 * D][:, pm]], 1).contiguous()
            w_q, b_q = w_kvq[:, 2 * D:].contiguous(), b_kvq[:, 2 * D:].contiguous()
            w_all, b_all = torch.cat([w_kv, w_q], 1).contiguous(), torch.cat([b_kv, b_q], 1).contiguous()
            # relation_pri[r, h] (the prior on the scores, :100) scales the rows of head h of the query transform
            pri = self.relation_pri.repeat_interleave(self.d_k, 1).unsqueeze(2)                                     # [R, D, 1]
            att = torch.stack([torch.block_diag(*self.relation_att[r]) for r in range(self.num_relations)]) * pri  # [R, D(n), D(k)]
            msg = torch.stack([torch.block_diag(*self.relation_msg[r]).t() for r in range(self.num_relations)])    # [R, D(n), D(k)]
            att, msg = att[:, pm, :].contiguous(), msg[:, :, pm].contiguous()
            return dict(w_kv=ops.to_operand(w_kv, opf), b_kv=b_kv, w_q=ops.to_operand(w_q, opf), b_q=b_q,
                        w_kvq=ops.to_operand(w_all, opf), b_kvq=b_all,
                        w_att=ops.to_operand(att, opf), w_msg=ops.to_operand(msg, opf), wa=ops.to_operand(wa, opf), ba=ba,
                        skip=skip, gamma=gamma, beta=beta)

        return self._packs.get(("tc", opf, tuple(order)), params, build)

    def _forward_packed_tc(self, plan: GraphPlan, x, order, grp, segs, opf, x_op=None, want_op: bool = False):
        """The layer on tensor cores end to end (models/HGT.py:68-127): every dense product is a tcgen05 grouped GEMM on
        operand-form inputs, every producer writes the operand form its consumer reads (no conversion passes between
        them), the edge kernel is the lane-grouped one of HEAT running over the relation-sorted (dst, relation) segments."""
        D, H = self.out_dim, self.n_heads
        pk = self._packed_tc(order, opf, x.device)
        tpc = plan.type_ptr_c()
        S = segs["S"]
        xs = x_op if x_op is not None else ops.to_operand(x, opf)
        st16 = opf != ops.OPF_BF16X3
        # (single-pass formats: q, q'_seg and the segment messages leave their GEMMs in the 16-bit storage form only - the
        #  [S, D] tensors are the bulk of the layer's HBM traffic; the 3-term format keeps them fp32)
        if opf == ops.OPF_BF16 and ops.tc_ok(plan.N, self.in_dim, 3 * D):
            # bf16 storage of K | V (BASELINE config 3): K | V | Q leave ONE GEMM as 16-bit rows, nothing else is written
            _, kvq = ops.typed_linear_op(xs, pk["w_kvq"], pk["b_kvq"], plan.type_ptr, 3 * D, want_y=False, want_op=True,
                                         type_ptr_c=tpc, opf=opf)
            kv, q32, q16 = kvq[:, :2 * D], None, kvq[:, 2 * D:]
        else:
            if opf == ops.OPF_BF16:
                _, kv = ops.typed_linear_op(xs, pk["w_kv"], pk["b_kv"], plan.type_ptr, 2 * D, want_y=False, want_op=True,
                                            type_ptr_c=tpc, opf=opf)
            else:
                kv, _ = ops.typed_linear_op(xs, pk["w_kv"], pk["b_kv"], plan.type_ptr, 2 * D, type_ptr_c=tpc, opf=opf)
            # q'_seg = relation_att[r, h] . q[dst, h] for the segments in relation order   (:88-92)
            q32, q16 = ops.typed_linear_op(xs, pk["w_q"], pk["b_q"], plan.type_ptr, D, type_ptr_c=tpc, opf=opf,
                                           want_y=not st16, want_op=st16)
        qg = ops.gather_rows16(q16, grp["dst_of_order"]) if st16 else ops.gather_to_operand(q32, grp["dst_of_order"], opf)
        qseg32, qseg16 = ops.typed_linear_op(qg, pk["w_att"], None, grp["rel_ptr"], D, type_ptr_c=grp["rel_ptr_c"], opf=opf,
                                             want_y=not st16, want_op=st16)
        qseg = qseg16 if st16 else qseg32
        sg = grp["seg_graph"]
        aggseg = ops.hetero_attn_work(kv[:, :D], kv[:, D:], qseg, sg["work"], sg["e_src"], sg["e_sim"], sg["e_rel"], sg["inv"],
                                      sg["ew"], sg["eb"], D, H, op_out=True, opf=opf)                       # :95-104
        msg32, msg16 = ops.typed_linear_op(aggseg, pk["w_msg"], None, grp["rel_ptr"], D, type_ptr_c=grp["rel_ptr_c"],
                                           opf=opf, want_y=not st16, want_op=st16)                          # :93
        msgseg = msg16 if st16 else msg32
        _, aggs = ops.segment_combine(msgseg, segs["row_seg_ptr"], plan.node_inv_r, plan.N, D, seg_pos=grp["seg_pos"],
                                      want_out=False, op_out=True, opf=opf)                                 # :105-106
        mask = None
        if self.training and self.drop.p > 0:
            mask = F.dropout(torch.ones((plan.N, D), dtype=torch.float32, device=x.device), self.drop.p, True)
        out, _ = ops.typed_linear_op(aggs, pk["wa"], pk["ba"], plan.type_ptr, D, skip=pk["skip"], res=x,
                                     row_gate=plan.node_inv_r, drop_mask=mask, type_ptr_c=tpc, opf=opf)     # :121-122
        if self.use_norm:
            if want_op:
                return ops.typed_layernorm(out, pk["gamma"], pk["beta"], plan.type_ptr, type_ptr_c=tpc, inplace=True,
                                           row_gate=plan.node_inv_r, op_out=True, opf=opf)
            out = _gated_layernorm(out, pk["gamma"], pk["beta"], plan, tpc)
        return out

    def forward_train(self, plan: GraphPlan, x: torch.Tensor) -> torch.Tensor:
        """Differentiable layer (autograd.py): what `loss.backward()` of the reference's train_one_step needs
        (trainer/train_gnn.py:68-71).  Same schedule as forward_packed; the edge attention runs in the lane-grouped
        column order (the backward kernel's layout), so K / V are projected with row-permuted weights and the transformed
        queries / aggregated segments are permuted on the way in / out."""
        D, H, dk = self.out_dim, self.n_heads, self.d_k
        perm = ops.head_perm(D, H)
        if perm is None:
            raise NotImplementedError(f"training needs the lane-grouped attention layout (D % 128 == 0, H a power of two "
                                      f"<= 32); got D={D}, H={H}")
        order = _graph_type_order(plan, self.node_dict)
        pm = perm.to(x.device)
        inv_pm = torch.empty_like(pm)
        inv_pm[pm] = torch.arange(D, device=x.device)
        tpc = plan.type_ptr_c()
        segs = plan.segments()
        grp = _relation_groups(plan, self.edge_dict, id(self.edge_dict))
        S = segs["S"]
        wk, bk = stack_linears(self.k_linears, order, row_perm=pm)
        wv, bv = stack_linears(self.v_linears, order, row_perm=pm)
        wq, bq = stack_linears(self.q_linears, order)
        wa, ba = stack_linears(self.a_linears, order)
        kvq = TypedLinearFn.apply(x, torch.cat([wk, wv, wq], 1), torch.cat([bk, bv, bq], 1), plan.type_ptr, tpc)
        k, v, q = kvq[:, :D], kvq[:, D:2 * D], kvq[:, 2 * D:]
        if S == 0:
            agg = torch.zeros_like(x)
        else:
            qseg = RelTransformFn.apply(q, self.relation_att, grp["dst_of_order"], grp["order"], grp["rel_ptr_c"],
                                        grp["rel_ptr"], grp["R"], H, dk, False, S, segs["row_seg_ptr"])      # :88-92
            pri = self.relation_pri.index_select(0, grp["seg_rel"].to(torch.int64))                           # [S, H]  :100
            qseg = (qseg.view(S, H, dk) * pri.unsqueeze(-1)).view(S, D)
            aggseg = SegAttnFn.apply(k, v, _Permute.apply(qseg, pm, 1).contiguous(), plan, D, H)              # :95-104
            aggseg = _Permute.apply(aggseg, inv_pm, 1).contiguous()
            msgseg = RelTransformFn.apply(aggseg, self.relation_msg, grp["order"], grp["order"], grp["rel_ptr_c"],
                                          grp["rel_ptr"], grp["R"], H, dk, True, S, None)                      # :93
            agg = SegmentCombineFn.apply(msgseg, plan, D)                                                      # :105-106
        lin = self.drop(TypedLinearFn.apply(agg, wa, ba, plan.type_ptr, tpc))                                  # :121
        skip_t = self.skip[torch.tensor(order, device=x.device)]
        out = SkipMixFn.apply(lin, x, skip_t, plan.type_ptr, plan.node_inv_r)                                  # :122, passthrough :118-120
        if self.use_norm:                                                                                      # :123-124
            parts = []
            for t, i in enumerate(order):
                a, b = plan.type_ptr[t], plan.type_ptr[t + 1]
                if b > a:
                    ln = F.layer_norm(out[a:b], (D,), self.norms[i].weight, self.norms[i].bias, self.norms[i].eps)
                    parts.append(torch.where(plan.node_inv_r[a:b].unsqueeze(1) != 0, ln, out[a:b]))
            out = torch.cat(parts, 0) if parts else out
        return out

    def forward(self, G: HeteroGraph, h: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        plan = G.plan()
        x = packed_features(G, plan, h)
        return unpack_rows(plan, self.forward_packed(plan, x))


def _gated_layernorm(x, gamma, beta, plan: GraphPlan, tpc):
    """LayerNorm on the rows that received a message; rows of a node type without incoming relation pass through
    (reference models/HGT.py:118-120 `continue`s before the norm).  Gated per ROW by 1/R (0 = passthrough): in a
    pack()ed batch "no incoming relation" is a property of the (type, graph) segment, not of the node type."""
    return ops.typed_layernorm(x, gamma, beta, plan.type_ptr, type_ptr_c=tpc, inplace=True, row_gate=plan.node_inv_r)


class HGT(nn.Module):
    """reference models/HGT.py:130-209.  forward(G, h=None) -> logits [B, out_dim]."""

    def __init__(self, node_dict, edge_dict, in_dim, hidden_dim, out_dim, n_layers, n_heads, use_norm=True,
                 graph_pooling_type="mean"):
        super().__init__()
        self.node_dict = node_dict
        self.edge_dict = edge_dict
        self.gcs = nn.ModuleList()
        self.in_dim = in_dim
        self.hidden_dim = hidden_dim
        self.out_dim = out_dim
        self.n_layers = n_layers
        self.graph_pooling_type = _check_pool(graph_pooling_type)
        self.adapt_ws = nn.ModuleList([nn.Linear(in_dim, hidden_dim) for _ in range(len(node_dict))])
        for _ in range(n_layers):
            self.gcs.append(HGTLayer(hidden_dim, hidden_dim, node_dict, edge_dict, n_heads, use_norm=use_norm))
        self.out = nn.Linear(hidden_dim, out_dim)                         # unused (models/HGT.py:150)
        self.linears_prediction = nn.ModuleDict({
            k: nn.ModuleList([nn.Linear(hidden_dim, out_dim) for _ in range(n_layers + 1)]) for k in node_dict})
        self._packs = PackCache()

    def prepare_plan(self, plan: GraphPlan):
        """Plan-side structures with host reads (segments, relation groups), built ahead by the streaming evaluator."""
        plan.segments()
        if self.gcs:
            _relation_groups(plan, self.edge_dict, id(self.edge_dict))

    def forward(self, G: HeteroGraph, h=None, return_embeddings: bool = False):
        plan = G.plan()
        T, B = len(plan.ntypes), plan.B
        order = _graph_type_order(plan, self.node_dict)
        names = list(plan.ntypes)
        x = packed_features(G, plan, h)
        if torch.is_grad_enabled() and any(p.requires_grad for p in param_list(self, "all", self.parameters)):
            return self._forward_train(G, plan, x, order, names, return_embeddings)
        params = param_list(self, "in", lambda: (p for m in self.adapt_ws for p in m.parameters()))
        w_in, b_in = self._packs.get(("in", tuple(order)), params, lambda: stack_linears(self.adapt_ws, order))
        x = ops.typed_linear(x, w_in, b_in, plan.type_ptr, act=ops.ACT_GELU, type_ptr_c=plan.type_ptr_c())  # :176-184
        scale = readout_scale(plan, G.independent)
        hg = None
        x_op = None
        n_run = self.n_layers if return_embeddings else self.n_layers
        for i in range(n_run):                                             # :189-199
            pp = param_list(self, ("pred", i, tuple(names)), lambda: (p for nt in names for p in self.linears_prediction[nt][i].parameters()))
            w_p, b_p = self._packs.get(("pred", i, tuple(names)), pp, lambda i=i: (
                torch.stack([self.linears_prediction[nt][i].weight for nt in names]).contiguous(),
                torch.stack([self.linears_prediction[nt][i].bias for nt in names]).contiguous()))
            if self.out_dim <= ops.AFFINE_MAX_OUT:        # fused pool + per-type prediction + sum over types / layers
                hg = ops.segment_pool_affine(x, plan.seg_ptr, T, B, self.graph_pooling_type, w_p, b_p, None, scale,
                                             out=hg, accumulate=hg is not None)
            else:
                pooled = ops.segment_pool(x, plan.seg_ptr, T * B, self.graph_pooling_type)
                o = ops.typed_linear(pooled, w_p, b_p, plan.readout_ptr(), row_scale=scale).view(T, B, -1).sum(0)
                hg = o if hg is None else hg + o
            # the output of the LAST layer is never read by the reference (models/HGT.py:199-209);
            # it is computed only when the caller asks for the embeddings
            if i + 1 < self.n_layers or return_embeddings:
                # (the layer after this one runs iff i + 2 < n_layers or embeddings are asked for: only then is the
                #  operand-form copy of h' wanted)
                nxt = i + 2 < self.n_layers or (return_embeddings and i + 1 < self.n_layers)
                x, x_op = self.gcs[i].forward_packed_pair(plan, x, x_op, want_op=nxt)
        if hg is None:
            hg = torch.zeros(B, self.out_dim, device=x.device)
        return (hg, unpack_rows(plan, x)) if return_embeddings else hg

    def _forward_train(self, G: HeteroGraph, plan: GraphPlan, x, order, names, return_embeddings: bool):
        """The differentiable chain (training): the same reads of the reference - logits = sum over layers i < L and node
        types of linears_prediction[type][i](pool(h^(i))) (models/HGT.py:189-199) - through autograd Functions."""
        T, B = len(plan.ntypes), plan.B
        w_in, b_in = stack_linears(self.adapt_ws, order)
        x = F.gelu(TypedLinearFn.apply(x, w_in, b_in, plan.type_ptr, plan.type_ptr_c()))                    # :176-184
        scale = readout_scale(plan, G.independent).unsqueeze(1)
        hg = None
        for i in range(self.n_layers):
            pooled = SegmentPoolFn.apply(x, plan, T * B, self.graph_pooling_type)
            w_p = torch.stack([self.linears_prediction[nt][i].weight for nt in names])
            b_p = torch.stack([self.linears_prediction[nt][i].bias for nt in names])
            o = (TypedLinearFn.apply(pooled, w_p, b_p, plan.readout_ptr(), None) * scale).view(T, B, -1).sum(0)
            hg = o if hg is None else hg + o
            if i + 1 < self.n_layers or return_embeddings:
                x = self.gcs[i].forward_train(plan, x)
        if hg is None:
            hg = torch.zeros(B, self.out_dim, device=x.device)
        return (hg, unpack_rows(plan, x)) if return_embeddings else hg
