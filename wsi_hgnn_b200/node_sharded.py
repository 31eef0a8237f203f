"""Node-sharded forward of ONE large slide over the GPUs of a box (BASELINE config 4:
100k-node TCGA-COAD-shape slide, k-NN graph_constructor + HEATNet4 forward on 4 x B200).

The reference runs a slide on one device (trainer/train_gnn.py:59-62); this is the same arithmetic
(models/HEATNet4.py:85-138, 195-247 / models/HEATNet2.py) with the dst rows of the relation-grouped CSR cut into
`world` contiguous ranges:

  * every rank owns the rows [r0, r1) of the type-major packed node order (ranges balanced on in-edges + rows),
    projects K|V|Q for its own rows only (typed GEMM on the local row segments of every node type),
  * ONE exchange step per layer: the K|V rows of all ranks are all-gathered into a padded [world, n_max, 2D] buffer
    (feature-space k-NN has no locality - the halo of a row range is ~ every row, SURVEY §8e - so an all-gather is
    the exchange that moves the fewest bytes: every row crosses NVLink once instead of once per referencing rank);
    the source ids of the local edges are remapped once to that padded layout, so the edge-attention kernel gathers
    from the gathered buffer exactly as it does from a local one (one 2*D*4-byte bulk copy per edge),
  * attention + a_linear epilogue run on the local dst rows; nothing else is exchanged per layer,
  * readout: per-rank partial (sum | max, count) of the (type, graph) segments -> one all-reduce of [T*B, D+1]
    -> the (tiny) prediction heads replicated on every rank.

The k-NN edge builder shards the same way (`knn_edges_sharded`): query rows are split over the ranks, candidates are
all rows; each rank owns complete neighbour lists, so there is no merge step - only an all-gather of the lists.

Collectives go through a small `Comm` interface: `DistComm` = torch.distributed (NCCL over NVLink on the GPUs;
gloo in the CPU tests of the partition / exchange plumbing), `LocalComm` = several virtual ranks inside one process
sharing one device (how the single-GPU parity test drives the very same rank code).
"""
from typing import Dict, List, Optional, Sequence

import torch

from . import ops
from .hetero_graph import GraphPlan, HeteroGraph


# ------------------------------------------------------------------------------------------------ partition
def balanced_row_ranges(rowptr_host: Sequence[int], world: int, row_cost: float = 8.0) -> List[int]:
    """Cut rows [0, N) into `world` contiguous ranges with (almost) equal cost, cost(row) = in_degree + row_cost
    (the edge phase scales with the in-edges, the projections with the rows).  -> boundaries [world + 1].
    Deterministic, O(N) on the host (one-time planning)."""
    if world < 1:
        raise ValueError("world must be >= 1")
    n = len(rowptr_host) - 1
    total = float(rowptr_host[n] - rowptr_host[0]) + row_cost * n
    bounds, r = [0], 0
    for p in range(1, world):
        target = total * p / world
        # first row index whose prefix cost reaches the target (prefix(i) = edges before row i + row_cost * i)
        lo, hi = r, n
        while lo < hi:
            mid = (lo + hi) // 2
            if (rowptr_host[mid] - rowptr_host[0]) + row_cost * mid < target:
                lo = mid + 1
            else:
                hi = mid
        r = lo
        bounds.append(r)
    bounds.append(n)
    return bounds


def clip_ptr(ptr: Sequence[int], r0: int, r1: int) -> List[int]:
    """Row-range pointers (type_ptr / seg_ptr) restricted to the rows [r0, r1), rebased to r0."""
    return [min(max(int(p), r0), r1) - r0 for p in ptr]


def padded_ids(ids: torch.Tensor, bounds: Sequence[int], n_max: int) -> torch.Tensor:
    """global packed row id -> row of the padded all-gather buffer [world * n_max]: owner * n_max + (id - r0[owner])."""
    b = torch.tensor(list(bounds), dtype=torch.int64, device=ids.device)
    owner = torch.bucketize(ids.to(torch.int64), b[1:-1], right=True)
    return (owner * n_max + ids.to(torch.int64) - b[owner]).to(torch.int32)


# ------------------------------------------------------------------------------------------------ collectives
class DistComm:
    """torch.distributed collectives (one process per GPU; NCCL over NVLink / NVSwitch, gloo on CPU)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def all_gather_blocks(self, buf: torch.Tensor, async_op: bool = False):
        """buf [world, n_max, C]: block `rank` holds this rank's rows; on completion every block is filled.
        async_op: return a work handle (wait() orders the current stream after the collective) so that the exchange
        overlaps whatever is enqueued in between - the Q projection in the node-sharded layer."""
        if self.world == 1:
            return None
        flat = buf.view(self.world, -1)
        try:
            w = self.dist.all_gather_into_tensor(flat.view(-1), flat[self.rank], group=self.group, async_op=async_op)
        except (RuntimeError, NotImplementedError):
            outs = [flat[r] for r in range(self.world)]
            mine = flat[self.rank].clone()
            w = self.dist.all_gather(outs, mine, group=self.group, async_op=async_op)
        return w if async_op else None

    def all_reduce(self, t: torch.Tensor, op: str = "sum"):
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM if op == "sum" else self.dist.ReduceOp.MAX, group=self.group)


class LocalComm:
    """`world` virtual ranks in ONE process (single-GPU parity test, debugging): the ranks are driven phase by phase
    by `run_virtual_ranks`; a collective is posted by every rank (k-th call of its kind within the phase) and completed
    by `_resolve_local` once the phase has run on all ranks."""

    def __init__(self, world: int):
        self.world = world
        self.pending: Dict[tuple, List] = {}
        self.calls: List[Dict[str, int]] = [dict() for _ in range(world)]

    def view(self, rank: int) -> "_LocalRankComm":
        return _LocalRankComm(self, rank)

    def post(self, rank: int, kind: str, buf: torch.Tensor):
        k = self.calls[rank].get(kind, 0)
        self.calls[rank][kind] = k + 1
        self.pending.setdefault((kind, k), [None] * self.world)[rank] = buf


class _LocalRankComm:
    def __init__(self, hub: LocalComm, rank: int):
        self.hub, self.rank, self.world = hub, rank, hub.world

    def all_gather_blocks(self, buf: torch.Tensor, async_op: bool = False):
        self.hub.post(self.rank, "gather", buf)
        return None

    def all_reduce(self, t: torch.Tensor, op: str = "sum"):
        self.hub.post(self.rank, "reduce_" + op, t)


def _resolve_local(hub: LocalComm):
    """Complete the collectives the virtual ranks posted in the phase that just ran."""
    pending, hub.pending = hub.pending, {}
    hub.calls = [dict() for _ in range(hub.world)]
    for (kind, _), bufs in pending.items():
        if any(b is None for b in bufs):
            raise RuntimeError(f"virtual rank missed collective {kind}")
        if kind == "gather":
            for r, src in enumerate(bufs):
                for d, dst in enumerate(bufs):
                    if d != r:
                        dst[r].copy_(src[r])
        else:
            stacked = torch.stack(bufs)
            red = stacked.sum(0) if kind == "reduce_sum" else stacked.max(0).values
            for b in bufs:
                b.copy_(red)


# ------------------------------------------------------------------------------------------------ the sharded forward
class NodeShardedHEAT:
    """Rank-local state + phases of the node-sharded HEATNet2 / HEATNet4 forward (inference).

        sh = NodeShardedHEAT(model, G, DistComm())      # G: the whole slide's structure; features of the own rows used
        logits = sh.forward()                           # identical on every rank

    Phases (what `forward` runs; `run_virtual_ranks` interleaves them over several virtual ranks):
        input_projection -> for every layer: project (posts the K|V all-gather) , aggregate -> pool (posts the
        all-reduce) -> finish.
    """

    def __init__(self, model, G: HeteroGraph, comm, row_cost: float = 8.0, bounds: Optional[Sequence[int]] = None,
                 kv_wire: str = "auto"):
        """kv_wire: storage of the exchanged K|V rows - "fp32", or "16" = the 16-bit type of the current matmul precision
        (fp16 under the default "fp16" precision, bf16 under "bf16"): the K|V GEMM writes the 16-bit rows straight into
        this rank's block of the gather buffer, the all-gather moves half the bytes and the edge kernel gathers half the
        bytes (fp32 scores / softmax / accumulation).  "auto" = "16" whenever the precision has a 16-bit operand type."""
        from .models.heat import _graph_type_order
        if model.training:
            raise NotImplementedError("NodeShardedHEAT is the inference path (model.eval())")
        self.model, self.G, self.comm = model, G, comm
        self.rank, self.world = comm.rank, comm.world
        plan: GraphPlan = G.plan()
        plan.check()
        self.plan = plan
        dev = plan.device
        rowptr_host = plan.rowptr.cpu().tolist()
        self.bounds = list(bounds) if bounds is not None else balanced_row_ranges(rowptr_host, self.world, row_cost)
        r0, r1 = self.bounds[self.rank], self.bounds[self.rank + 1]
        self.r0, self.r1, self.n_loc = r0, r1, r1 - r0
        self.n_max = max(1, max(self.bounds[p + 1] - self.bounds[p] for p in range(self.world)))
        e0, e1 = rowptr_host[r0], rowptr_host[r1]
        self.order = _graph_type_order(plan, model.node_dict)
        self.type_ptr = clip_ptr(plan.type_ptr, r0, r1)
        self.seg_ptr_host = clip_ptr(plan.seg_ptr_host, r0, r1)
        self.seg_ptr = torch.tensor(self.seg_ptr_host, dtype=torch.int32, device=dev)
        # local CSR: own dst rows, source ids in the padded all-gather layout
        self.rowptr = (plan.rowptr[r0:r1 + 1] - e0).contiguous()
        self.e_src = padded_ids(plan.e_src[e0:e1], self.bounds, self.n_max).contiguous()
        self.e_sim = plan.e_sim[e0:e1].contiguous()
        self.e_rel = plan.e_rel[e0:e1].contiguous()
        self.inv_r = plan.node_inv_r[r0:r1].contiguous()
        self.work = ops.plan_attn_work(self.rowptr, self.e_rel, self.n_loc, 16) if self.n_loc > 0 else None
        D = model.gcs[0].out_size if len(model.gcs) else model.adapt_ws[0].out_features
        self.D = D
        if kv_wire not in ("auto", "fp32", "16"):
            raise ValueError("kv_wire must be 'auto', 'fp32' or '16'")
        self.opf = ops.matmul_opf()
        F_in = model.adapt_ws[0].in_features
        # the tensor-core operand chain (operands converted once, 16-bit copies emitted by the producing kernels) needs
        # tile-friendly local shapes; otherwise the fp32-API GEMMs (per-call conversion) are used
        # (decided from the row ranges of ALL ranks - every rank knows them - so that the ranks agree on the wire type)
        def chain_ok(n):
            return (n >= 512 and ops.tc_ok(n, F_in, D) and ops.tc_ok(n, D, 2 * D) and ops.tc_ok(n, D, D))
        all_chain = (len(model.gcs) > 0 and ops.head_perm(D, model.gcs[0].n_heads) is not None and
                     all(chain_ok(self.bounds[p + 1] - self.bounds[p]) for p in range(self.world)))
        self.chain = all_chain
        wire16 = kv_wire != "fp32" and self.opf in (ops.OPF_F16, ops.OPF_BF16) and all_chain
        if kv_wire == "16" and not wire16:
            raise NotImplementedError("kv_wire='16' needs the fp16 / bf16 matmul precision and tile-friendly shards on every rank")
        self.kv_dtype = (torch.float16 if self.opf == ops.OPF_F16 else torch.bfloat16) if wire16 else torch.float32
        self.kv_all = torch.zeros((self.world, self.n_max, 2 * D), dtype=self.kv_dtype, device=dev)
        self.x: Optional[torch.Tensor] = None
        self.x_op: Optional[torch.Tensor] = None
        self.q: Optional[torch.Tensor] = None
        self._gather = None
        self.pool_buf: Optional[torch.Tensor] = None
        self.halo_bytes_per_layer = (self.world - 1) * self.n_max * 2 * D * self.kv_all.element_size()   # received per rank per layer

    # -- phases ----------------------------------------------------------------------------------
    def input_projection(self, feat_local: Optional[torch.Tensor] = None):
        """x = adapt_ws[type](feat) for the own rows (models/HEATNet4.py:198-206)."""
        from .models.heat import packed_features
        from .models._packing import stack_linears
        m = self.model
        if feat_local is None:
            feat_local = packed_features(self.G, self.plan, None)[self.r0:self.r1]
        w_in, b_in = stack_linears(m.adapt_ws, self.order)
        if self.chain:
            w_in_op = m._packs.get(("ns_in_op", self.opf, tuple(self.order)), list(m.adapt_ws.parameters()),
                                   lambda: ops.to_operand(w_in.detach(), self.opf))
            self.x, self.x_op = ops.typed_linear_op(ops.to_operand(feat_local.contiguous().float(), self.opf), w_in_op,
                                                    b_in.detach(), self.type_ptr, int(w_in.shape[1]), want_op=True, opf=self.opf)
            return
        self.x = (ops.typed_linear(feat_local.contiguous(), w_in.detach(), b_in.detach(), self.type_ptr)
                  if self.n_loc > 0 else feat_local.new_zeros((0, w_in.shape[1])))

    def project(self, l: int):
        """K|V of the own rows straight into this rank's block of the gather buffer, the all-gather posted
        asynchronously, then Q of the own rows while the K|V rows travel."""
        layer = self.model.gcs[l]
        w_kvq, b_kvq, _, _, _, use_perm = layer._packed(self.order)
        if not use_perm:
            raise NotImplementedError("node-sharded forward needs the lane-grouped attention layout "
                                      "(D % 128 == 0, H a power of two <= 32)")
        D = self.D
        w_kv, b_kv, w_q, b_q = layer._packs.get(("kv|q", tuple(self.order)), [w_kvq, b_kvq], lambda: (
            w_kvq[:, :2 * D].contiguous(), b_kvq[:, :2 * D].contiguous(), w_kvq[:, 2 * D:].contiguous(),
            b_kvq[:, 2 * D:].contiguous()))
        if self.chain:
            w_kv_op, w_q_op = layer._packs.get(("ns_kv|q_op", self.opf, tuple(self.order)), [w_kvq, b_kvq], lambda: (
                ops.to_operand(w_kv, self.opf), ops.to_operand(w_q, self.opf)))
            mine = self.kv_all[self.rank, :self.n_loc]
            if self.kv_dtype == torch.float32:
                ops.typed_linear_op(self.x_op, w_kv_op, b_kv, self.type_ptr, 2 * D, out=mine, opf=self.opf)
            else:                                    # the 16-bit rows straight into this rank's block of the gather buffer
                ops.typed_linear_op(self.x_op, w_kv_op, b_kv, self.type_ptr, 2 * D, want_y=False, want_op=True, out_op=mine,
                                    opf=self.opf)
            self._gather = self.comm.all_gather_blocks(self.kv_all, async_op=True)
            self.q, _ = ops.typed_linear_op(self.x_op, w_q_op, b_q, self.type_ptr, D, opf=self.opf)
            return
        if self.n_loc > 0:
            ops.typed_linear(self.x, w_kv, b_kv, self.type_ptr, out=self.kv_all[self.rank, :self.n_loc])
        self._gather = self.comm.all_gather_blocks(self.kv_all, async_op=True)
        if self.n_loc > 0:
            self.q = ops.typed_linear(self.x, w_q, b_q, self.type_ptr)

    def aggregate(self, l: int):
        """edge attention over the own dst rows (sources from the gathered K|V) + a_linear / skip epilogue."""
        layer = self.model.gcs[l]
        _, _, wa, ba, skip, _ = layer._packed(self.order)
        D, H = self.D, layer.n_heads
        if self._gather is not None:
            self._gather.wait()
            self._gather = None
        if self.n_loc == 0:
            return
        kv = self.kv_all.view(self.world * self.n_max, 2 * D)
        if self.chain:
            wa_op = layer._packs.get(("ns_a_op", self.opf, tuple(self.order)), [wa], lambda: ops.to_operand(wa, self.opf))
            agg_op = ops.hetero_attn_work(kv[:, :D], kv[:, D:], self.q, self.work, self.e_src, self.e_sim, self.e_rel,
                                          self.inv_r, layer.e_linear.weight, layer.e_linear.bias, D, H, op_out=True,
                                          opf=self.opf)
            self.x, self.x_op = ops.typed_linear_op(agg_op, wa_op, ba, self.type_ptr, D, skip=skip, res=self.x,
                                                    row_gate=self.inv_r, want_op=l + 1 < len(self.model.gcs), opf=self.opf)
            return
        agg = ops.hetero_attn_work(kv[:, :D], kv[:, D:], self.q, self.work, self.e_src, self.e_sim,
                                   self.e_rel, self.inv_r, layer.e_linear.weight, layer.e_linear.bias, D, H)
        self.x = ops.typed_linear(agg, wa, ba, self.type_ptr, skip=skip, res=self.x, row_gate=self.inv_r)

    def pool(self):
        """partial typed readout of the own rows: [T*B, D + 1] = (sum or max | row count); posts the all-reduce(s)."""
        op = self.model.graph_pooling_type
        n_seg = len(self.seg_ptr_host) - 1
        D = self.D
        cnt = torch.tensor([self.seg_ptr_host[i + 1] - self.seg_ptr_host[i] for i in range(n_seg)],
                           dtype=torch.float32, device=self.kv_all.device)
        if self.n_loc > 0:
            part = ops.segment_pool(self.x, self.seg_ptr, n_seg, "max" if op == "max" else "sum")
        else:
            part = torch.zeros((n_seg, D), dtype=torch.float32, device=self.kv_all.device)
        self.pool_cnt = cnt
        if op == "max":
            part = torch.where(cnt.unsqueeze(1) > 0, part, torch.full_like(part, float("-inf")))
            self.pool_buf = part.contiguous()
            self.comm.all_reduce(self.pool_buf, "max")
            self.comm.all_reduce(self.pool_cnt, "sum")
        else:
            self.pool_buf = torch.cat([part, cnt.unsqueeze(1)], 1).contiguous()
            self.comm.all_reduce(self.pool_buf, "sum")

    def finish(self) -> torch.Tensor:
        """pooled [T*B, D] -> logits [B, out_dim] by the model's own prediction heads (replicated, tiny)."""
        op = self.model.graph_pooling_type
        D = self.D
        if op == "max":
            cnt = self.pool_cnt
            pooled = torch.where(cnt.unsqueeze(1) > 0, self.pool_buf, torch.zeros_like(self.pool_buf))
        else:
            cnt = self.pool_buf[:, D]
            pooled = self.pool_buf[:, :D]
            if op == "mean":
                pooled = pooled / cnt.clamp_min(1.0).unsqueeze(1)
        return self.model.logits_from_pooled(self.G, self.plan, pooled.contiguous())

    def forward(self, feat_local: Optional[torch.Tensor] = None) -> torch.Tensor:
        with torch.no_grad():
            self.input_projection(feat_local)
            for l in range(len(self.model.gcs)):
                self.project(l)
                self.aggregate(l)
            self.pool()
            return self.finish()

    def embeddings(self) -> torch.Tensor:
        """the own rows of the final node embeddings [n_loc, D] (after forward())."""
        return self.x


def run_virtual_ranks(model, G: HeteroGraph, world: int, row_cost: float = 8.0, bounds=None, kv_wire: str = "auto"):
    """Drive `world` virtual ranks of the node-sharded forward in one process on one device (LocalComm).
    -> (logits of rank 0, [per-rank logits], [per-rank embeddings], ranks)."""
    hub = LocalComm(world)
    ranks = [NodeShardedHEAT(model, G, hub.view(r), row_cost, bounds, kv_wire) for r in range(world)]
    with torch.no_grad():
        for s in ranks:
            s.input_projection()
        for l in range(len(model.gcs)):
            for s in ranks:
                s.project(l)
            _resolve_local(hub)
            for s in ranks:
                s.aggregate(l)
        for s in ranks:
            s.pool()
        _resolve_local(hub)
        outs = [s.finish() for s in ranks]
    return outs[0], outs, [s.embeddings() for s in ranks], ranks


# ------------------------------------------------------------------------------------------------ sharded edge builder
def knn_edges_rank(features: torch.Tensor, radius: int, q0: int, q1: int):
    """The query rows [q0, q1) of the edge builder: (nbr int32 [q1-q0, radius-1], sim fp32 [q1-q0, radius-1])."""
    k = radius - 1
    nbr = ops.knn_topk(features, radius, q0, q1)[:, 1:].contiguous()                         # rank 0 dropped (:270)
    src = torch.arange(q0, q1, device=features.device, dtype=torch.int64).repeat_interleave(k)
    sim, _ = ops.edge_pearson(features, src, nbr.reshape(-1).to(torch.int64))
    return nbr, sim.view(q1 - q0, k)


def knn_edges_sharded(features: torch.Tensor, radius: int, comm, bounds: Optional[Sequence[int]] = None):
    """construct_graph()'s edge arrays (construct_graph/graph_constructor.py:262-282) with the QUERY rows sharded over
    the ranks: rank p computes the `radius` nearest rows (exact, (distance, index) order, self included) and the
    Pearson similarity for its own queries against ALL rows, then the per-rank lists are all-gathered.
    features: the full [N, F] fp32 matrix on every rank (all-gather the row blocks first when they arrive sharded).
    -> (edge_index int64 [2, N*(radius-1)], edge_type uint8 [E], sim fp32 [E]) identical to the single-GPU builder."""
    n = int(features.shape[0])
    world, rank = comm.world, comm.rank
    if bounds is None:
        bounds = [n * p // world for p in range(world + 1)]
    q0, q1 = bounds[rank], bounds[rank + 1]
    k = radius - 1
    n_max = max(1, max(bounds[p + 1] - bounds[p] for p in range(world)))
    dev = features.device
    nbr_all = torch.zeros((world, n_max, k), dtype=torch.int32, device=dev)
    sim_all = torch.zeros((world, n_max, k), dtype=torch.float32, device=dev)
    if q1 > q0:
        nbr_all[rank, :q1 - q0], sim_all[rank, :q1 - q0] = knn_edges_rank(features, radius, q0, q1)
    comm.all_gather_blocks(nbr_all)
    comm.all_gather_blocks(sim_all)
    return _assemble_edges(nbr_all, sim_all, bounds, k)


def _assemble_edges(nbr_all, sim_all, bounds, k):
    world = nbr_all.shape[0]
    dev = nbr_all.device
    nbr = torch.cat([nbr_all[p, :bounds[p + 1] - bounds[p]] for p in range(world)], 0)
    sim = torch.cat([sim_all[p, :bounds[p + 1] - bounds[p]] for p in range(world)], 0).reshape(-1)
    n = nbr.shape[0]
    src = torch.arange(n, device=dev, dtype=torch.int64).repeat_interleave(k)
    dst = nbr.reshape(-1).to(torch.int64)
    return torch.stack([src, dst]), (sim > 0).to(torch.uint8), sim
