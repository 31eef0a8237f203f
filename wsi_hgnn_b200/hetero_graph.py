"""HeteroGraph: the graph container the hot path runs on (no DGL dependency).

It exposes exactly the subset of the ``DGLHeteroGraph`` surface that the
reference models and their callers touch:

* ``G.ntypes`` / ``G.canonical_etypes`` / ``G.etypes``        (reference models/HEATNet4.py:91,122,199)
* ``G.nodes[nt].data['feat']``                                 (models/HEATNet4.py:202)
* ``G.edata['sim']`` -> dict keyed by canonical etype           (models/HEATNet4.py:209-210)
* ``G[s, e, d]`` relation view, ``G.local_scope()``, ``G.to()`` (models/HEATNet4.py:92,228; trainer/train_gnn.py:60,64)
* ``G.is_homogeneous``                                          (data.py:120)
* ``batch_size`` / ``batch_num_nodes(ntype)``                   (what dgl.readout.*_nodes consume, pooling/avg_pooling.py:15-17)

Two ways of putting several slides into one object:

* :func:`batch` - DGL ``dgl.batch`` semantics: all graphs must share the same
  relation set; the cross-relation mean denominator R_t is that of the shared
  metagraph.
* :func:`pack`  - the reference trainer's *tuple* branch
  (trainer/train_gnn.py:59-62, ``torch.cat([gnn(g) for g in graphs])``):
  graphs are laid out block-diagonally but keep their own relation sets, their
  own R_t and their own "type has no nodes" readout behaviour, so one launch
  over the packed object equals the concatenation of independent forwards.

The device-side layout the CUDA kernels consume is produced by
:meth:`HeteroGraph.plan` (see ``GraphPlan``).
"""
from __future__ import annotations

import contextlib
from typing import Dict, List, Optional, Sequence, Tuple

import torch

CEType = Tuple[str, str, str]


class _Frame(dict):
    """A per-type (or per-relation) feature dictionary."""


class _NodeTypeView:
    def __init__(self, graph: "HeteroGraph", ntype: str):
        self._g = graph
        self._nt = ntype

    @property
    def data(self) -> _Frame:
        return self._g._ndata[self._nt]


class _NodesAccessor:
    def __init__(self, graph: "HeteroGraph"):
        self._g = graph

    def __getitem__(self, ntype: str) -> _NodeTypeView:
        if ntype not in self._g._ndata:
            raise KeyError(ntype)
        return _NodeTypeView(self._g, ntype)


class _TypedDataView:
    """``G.ndata`` / ``G.edata``: name -> tensor (one type) or {type: tensor}."""

    def __init__(self, frames: Dict, keys: Sequence):
        self._frames = frames
        self._keys = list(keys)

    def __getitem__(self, name: str):
        if len(self._keys) == 1:
            return self._frames[self._keys[0]][name]
        out = {k: self._frames[k][name] for k in self._keys if name in self._frames[k]}
        if not out:
            raise KeyError(name)
        return out

    def __setitem__(self, name: str, value):
        if isinstance(value, dict):
            for k, v in value.items():
                self._frames[k][name] = v
        else:
            if len(self._keys) != 1:
                raise ValueError("a dict keyed by type is required when there is more than one type")
            self._frames[self._keys[0]][name] = value

    def __contains__(self, name: str) -> bool:
        return any(name in self._frames[k] for k in self._keys)

    def pop(self, name: str):
        out = self[name]
        for k in self._keys:
            self._frames[k].pop(name, None)
        return out

    def update(self, d: Dict):
        for k, v in d.items():
            self[k] = v


class RelationView:
    """``G[s, e, d]``: one relation of the parent graph (shares its frames)."""

    def __init__(self, graph: "HeteroGraph", cetype: CEType):
        self._g = graph
        self.cetype = cetype

    @property
    def canonical_etypes(self):
        return [self.cetype]

    def edges(self):
        return self._g._edges[self.cetype]

    def num_edges(self) -> int:
        return int(self._g._edges[self.cetype][0].shape[0])

    @property
    def edata(self) -> _Frame:
        return self._g._edata[self.cetype]

    @property
    def srcdata(self) -> _Frame:
        return self._g._ndata[self.cetype[0]]

    @property
    def dstdata(self) -> _Frame:
        return self._g._ndata[self.cetype[2]]

    def num_src_nodes(self) -> int:
        return self._g.num_nodes(self.cetype[0])

    def num_dst_nodes(self) -> int:
        return self._g.num_nodes(self.cetype[2])


def _to_device_async(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    """small host array -> device without a stream synchronisation (pinned staging + non-blocking copy on CUDA)."""
    if dev.type != "cuda":
        return t.to(dev)
    return t.pin_memory().to(dev, non_blocking=True)


class GraphPlan:
    """Device-resident layout of one (possibly batched/packed) graph.

    Nodes are *packed type-major*: packed id = type_ptr[t] + local id, so every
    per-type feature matrix is a contiguous row range of one [N, D] buffer and
    a typed linear is a grouped GEMM over row segments.

    Edges are sorted by (dst packed id, relation index, original edge id); one
    dst row of the CSR therefore holds its in-edges grouped by relation, which
    is what the per-(dst, relation, head) softmax needs
    (reference models/HEATNet4.py:109-119).

    Tensors (all on the graph's device):
      type_ptr_dev int32 [T+1]     row range of each node type
      seg_ptr      int32 [T*B+1]   (type, graph) readout segments, type-major
      rowptr       int32 [N+1]     CSR over packed dst ids
      e_src        int32 [E]       packed src id per edge (dst-sorted order)
      e_sim        fp32  [E]       edge attribute 'sim' (0 if the graph has none)
      e_rel        uint8 [E]       relation slot of the edge (index into rel_list)
      node_inv_r   fp32  [N]       1/R_t for the node's (graph,) type; 0 => no incoming relation (passthrough)
      t_rowptr     int32 [N+1]     transposed (src-major) CSR, built lazily for backward
      t_eid        int32 [E]       position in dst-sorted order of each src-sorted edge
      e_dst        int32 [E]       packed dst id per edge (dst-sorted order), lazily for backward
    """

    def __init__(self):
        self.ntypes: List[str] = []
        self.rel_list: List[CEType] = []
        self.type_ptr: List[int] = []
        self.N = 0
        self.E = 0
        self.B = 1
        self.type_ptr_dev = None
        self.seg_ptr = None
        self.seg_ptr_host: List[int] = []
        self.rowptr = None
        self.e_src = None
        self.e_sim = None
        self.e_rel = None
        self.node_inv_r = None
        self.rel_src_type: List[int] = []
        self.rel_dst_type: List[int] = []
        self.r_count: List[int] = []          # R_t per type (batch mode)
        self._max_in_degree = 0
        self._stats = None                    # device int32 [4] of the native builder, read lazily
        self._rel_table = None                # device int32 [3, R + 1]: edge range / src offset / dst offset per relation
        self._t = None
        self.device = torch.device("cpu")
        self.seg_nonempty = None              # bool [T, B] on host
        self.cache: Dict = {}                 # per-model derived tensors (relation id maps, packed feats, ...)
        self._segs = None

    def check(self):
        """Raise if the native builder flagged an out-of-range edge endpoint (reads 8 bytes back: a host sync,
        so it runs with the first work-list build rather than inside plan())."""
        if self._stats is not None:
            mx, bad = self._stats[:2].tolist()
            self._stats = None
            self._max_in_degree = mx
            if bad:
                raise IndexError("edge endpoint out of range")

    @property
    def max_in_degree(self) -> int:
        self.check()
        return self._max_in_degree

    @max_in_degree.setter
    def max_in_degree(self, v: int):
        self._max_in_degree = int(v)

    def type_ptr_c(self):
        """type_ptr as a ctypes int32 array (host argument of the typed kernels)."""
        if "type_ptr_c" not in self.cache:
            import ctypes
            self.cache["type_ptr_c"] = (ctypes.c_int32 * len(self.type_ptr))(*self.type_ptr)
        return self.cache["type_ptr_c"]

    def readout_ptr(self) -> List[int]:
        """Row ranges of the pooled [T*B, D] matrix by node type (type-major)."""
        return [t * self.B for t in range(len(self.ntypes) + 1)]

    def segments(self):
        """(dst, relation) segments of the dst-sorted edge array, for the HGT path.

        Returns dict: S, seg_ptr int32 [S+1], seg_dst int32 [S], seg_slot int64 [S] (graph relation
        slot), row_seg_ptr int32 [N+1] (segments of each dst row are contiguous)."""
        if self._segs is None:
            self.check()
            dev = self.device
            E = self.E
            if E == 0:
                z32 = torch.zeros(0, dtype=torch.int32, device=dev)
                self._segs = dict(S=0, seg_ptr=torch.zeros(1, dtype=torch.int32, device=dev), seg_dst=z32,
                                  seg_slot=torch.zeros(0, dtype=torch.int64, device=dev),
                                  row_seg_ptr=torch.zeros(self.N + 1, dtype=torch.int32, device=dev))
                return self._segs
            deg = (self.rowptr[1:] - self.rowptr[:-1]).to(torch.int64)
            e_dst = torch.repeat_interleave(torch.arange(self.N, device=dev, dtype=torch.int64), deg)
            rel = self.e_rel.to(torch.int64)
            key = e_dst * 256 + rel
            new = torch.ones(E, dtype=torch.bool, device=dev)
            new[1:] = key[1:] != key[:-1]
            starts = torch.nonzero(new).reshape(-1)
            S = int(starts.numel())
            seg_ptr = torch.empty(S + 1, dtype=torch.int32, device=dev)
            seg_ptr[:S] = starts.to(torch.int32)
            seg_ptr[S] = E
            seg_dst = e_dst[starts]
            counts = torch.bincount(seg_dst, minlength=self.N)
            row_seg_ptr = torch.zeros(self.N + 1, dtype=torch.int32, device=dev)
            row_seg_ptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
            self._segs = dict(S=S, seg_ptr=seg_ptr, seg_dst=seg_dst.to(torch.int32).contiguous(),
                              seg_slot=rel[starts].contiguous(), row_seg_ptr=row_seg_ptr)
        return self._segs

    def row_segments(self):
        """(segment index of every packed row int64 [N], 1 / rows-in-segment fp32 [N]) - backward of the typed readout."""
        if "row_segments" not in self.cache:
            lens = (self.seg_ptr[1:] - self.seg_ptr[:-1]).to(torch.int64)
            seg = torch.repeat_interleave(torch.arange(lens.numel(), device=self.device), lens)
            inv = (1.0 / lens.clamp_min(1).to(torch.float32)).index_select(0, seg)
            self.cache["row_segments"] = (seg, inv)
        return self.cache["row_segments"]

    def row_types(self):
        """node-type index of every packed row, int64 [N]."""
        if "row_types" not in self.cache:
            counts = torch.tensor([self.type_ptr[t + 1] - self.type_ptr[t] for t in range(len(self.ntypes))],
                                  device=self.device)
            self.cache["row_types"] = torch.repeat_interleave(torch.arange(len(self.ntypes), device=self.device), counts)
        return self.cache["row_types"]

    def attn_work(self, chunk: int = 16):
        """Work list of the edge-attention kernel (wsi_hetero_attn_work_fwd): rows with at most `chunk`
        in-edges are one item; rows with more (k-NN hubs) are cut, per (row, relation) segment, into chunks
        of <= `chunk` edges whose partials a merge launch combines.  The chunk items come first, then the
        whole rows sorted by edge count, largest first (the kernel deals items to warps round-robin).  dict: items int32 [n_items, 4], n_items, split_row, split_ptr, part_rel, n_split, n_part."""
        key = ("attn_work", chunk)
        if key in self.cache:
            return self.cache[key]
        dev = self.device
        if dev.type == "cuda":
            from . import ops
            begun = self.cache.pop(("attn_work_begun", chunk), None)
            if begun is None:
                begun = ops.plan_attn_work_begin(self.rowptr, self.e_rel, self.N, chunk, self._stats)
            work = ops.plan_attn_work_finish(begun)
            if self._stats is not None:
                self._stats = None
                self._max_in_degree = work["max_in_degree"]
                if work["bad_edges"]:
                    raise IndexError("edge endpoint out of range")
            self.cache[key] = work
            return work
        i32 = dict(dtype=torch.int32, device=dev)
        rowptr = self.rowptr.to(torch.int64)
        deg = rowptr[1:] - rowptr[:-1]
        split = deg > chunk
        rows_c = torch.nonzero(~split).reshape(-1)
        rows_c = rows_c[torch.argsort(deg[rows_c], descending=True, stable=True)]     # largest first
        items_c = torch.stack([rows_c, rowptr[rows_c], rowptr[rows_c + 1], torch.full_like(rows_c, -1)], 1)
        n_split = int(split.sum()) if self.N > 0 else 0
        if n_split == 0:
            work = dict(items=items_c.to(torch.int32).contiguous(), n_items=int(items_c.shape[0]),
                        split_row=torch.zeros(1, **i32), split_ptr=torch.zeros(2, **i32),
                        part_rel=torch.zeros(1, **i32), n_split=0, n_part=0, sched=torch.zeros(2, **i32))
            self.cache[key] = work
            return work
        segs = self.segments()
        seg_ptr = segs["seg_ptr"].to(torch.int64)
        seg_dst = segs["seg_dst"].to(torch.int64)
        sel = torch.nonzero(split[seg_dst]).reshape(-1)              # segments of the split rows, in edge order
        s_beg, s_end, s_dst = seg_ptr[sel], seg_ptr[sel + 1], seg_dst[sel]
        n_ch = (s_end - s_beg + chunk - 1) // chunk
        seg_of = torch.repeat_interleave(torch.arange(sel.numel(), device=dev), n_ch)
        first = torch.cumsum(n_ch, 0) - n_ch
        c_idx = torch.arange(seg_of.numel(), device=dev) - first[seg_of]
        p_beg = s_beg[seg_of] + c_idx * chunk
        p_end = torch.minimum(p_beg + chunk, s_end[seg_of])
        n_part = int(seg_of.numel())
        slots = torch.arange(n_part, device=dev)
        items_p = torch.stack([s_dst[seg_of], p_beg, p_end, slots], 1)
        part_rel = segs["seg_slot"][sel][seg_of].to(torch.int32).contiguous()
        split_row = torch.nonzero(split).reshape(-1)
        per_row = torch.zeros(self.N, dtype=torch.int64, device=dev).index_add_(0, s_dst, n_ch)
        split_ptr = torch.zeros(n_split + 1, dtype=torch.int64, device=dev)
        split_ptr[1:] = torch.cumsum(per_row[split_row], 0)
        items = torch.cat([items_p, items_c], 0).to(torch.int32).contiguous()
        part_split = torch.repeat_interleave(torch.arange(n_split, device=dev), per_row[split_row])
        work = dict(items=items, n_items=int(items.shape[0]), split_row=split_row.to(torch.int32).contiguous(),
                    split_ptr=split_ptr.to(torch.int32).contiguous(), part_rel=part_rel, n_split=n_split,
                    n_part=n_part, part_split=part_split.to(torch.int32).contiguous(),
                    split_cnt=torch.zeros(n_split, dtype=torch.int32, device=dev),
                    sched=torch.zeros(2, dtype=torch.int32, device=dev))
        self.cache[key] = work
        return work

    def rows_by_degree(self):
        """int32 [N]: dst rows in decreasing in-degree order (ties by row id) - the processing order of the attention
        backward, whose one-row-per-warp blocks are dealt out by the hardware scheduler (heaviest rows first)."""
        if "rows_by_degree" not in self.cache:
            deg = (self.rowptr[1:] - self.rowptr[:-1]).to(torch.int64)
            self.cache["rows_by_degree"] = torch.argsort(deg, descending=True, stable=True).to(torch.int32).contiguous()
        return self.cache["rows_by_degree"]

    def attn_work_begin(self, chunk: int = 16):
        """Enqueue the counting half of the work-list build (no host wait); attn_work() completes it."""
        if self.device.type == "cuda" and self.E > 0 and ("attn_work", chunk) not in self.cache and \
                ("attn_work_begun", chunk) not in self.cache:
            from . import ops
            self.cache[("attn_work_begun", chunk)] = ops.plan_attn_work_begin(self.rowptr, self.e_rel, self.N, chunk,
                                                                               self._stats)

    def transposed(self):
        """(t_rowptr, t_eid, e_dst) for the backward scatter to src rows."""
        if self._t is None:
            dev = self.device
            deg = (self.rowptr[1:] - self.rowptr[:-1]).to(torch.int64)
            e_dst = torch.repeat_interleave(torch.arange(self.N, device=dev, dtype=torch.int32), deg)
            src64 = self.e_src.to(torch.int64)
            order = torch.argsort(src64, stable=True)
            counts = torch.bincount(src64, minlength=self.N)
            t_rowptr = torch.zeros(self.N + 1, dtype=torch.int32, device=dev)
            t_rowptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
            self._t = (t_rowptr, order.to(torch.int32).contiguous(), e_dst.contiguous())
        return self._t


class HeteroGraph:
    def __init__(self,
                 num_nodes_dict: Dict[str, int],
                 edges: Dict[CEType, Tuple[torch.Tensor, torch.Tensor]],
                 ndata: Optional[Dict[str, Dict[str, torch.Tensor]]] = None,
                 edata: Optional[Dict[CEType, Dict[str, torch.Tensor]]] = None):
        # [DGL-mem] dgl.heterograph sorts node type names and relation tuples.
        self.ntypes: List[str] = sorted(num_nodes_dict.keys())
        self._num_nodes = {k: int(num_nodes_dict[k]) for k in self.ntypes}
        self.canonical_etypes: List[CEType] = sorted(edges.keys())
        self._edges: Dict[CEType, Tuple[torch.Tensor, torch.Tensor]] = {}
        for ce in self.canonical_etypes:
            s, d = edges[ce]
            s = torch.as_tensor(s, dtype=torch.int64)
            d = torch.as_tensor(d, dtype=torch.int64)
            if s.shape != d.shape or s.dim() != 1:
                raise ValueError(f"relation {ce}: src/dst must be 1-D of equal length")
            if ce[0] not in self._num_nodes or ce[2] not in self._num_nodes:
                raise KeyError(f"relation {ce} names an unknown node type")
            self._edges[ce] = (s, d)
        self._ndata: Dict[str, _Frame] = {nt: _Frame() for nt in self.ntypes}
        self._edata: Dict[CEType, _Frame] = {ce: _Frame() for ce in self.canonical_etypes}
        if ndata:
            for nt, fr in ndata.items():
                for k, v in fr.items():
                    if v.shape[0] != self._num_nodes[nt]:
                        raise ValueError(f"ndata[{nt}][{k}] has {v.shape[0]} rows, expected {self._num_nodes[nt]}")
                    self._ndata[nt][k] = v
        if edata:
            for ce, fr in edata.items():
                for k, v in fr.items():
                    if v.shape[0] != self._edges[ce][0].shape[0]:
                        raise ValueError(f"edata[{ce}][{k}] has wrong length")
                    self._edata[ce][k] = v
        # batching info: one graph by default
        self._batch_num_nodes: Dict[str, List[int]] = {nt: [self._num_nodes[nt]] for nt in self.ntypes}
        self._batch_num_edges: Dict[CEType, List[int]] = {
            ce: [int(self._edges[ce][0].shape[0])] for ce in self.canonical_etypes}
        self.batch_size = 1
        # pack() mode: per-graph relation presence [B][R] (None => DGL batch semantics)
        self._rel_present: Optional[List[List[bool]]] = None
        self._plan: Optional[GraphPlan] = None

    # ------------------------------------------------------------------ basic queries
    @property
    def etypes(self) -> List[str]:
        return [ce[1] for ce in self.canonical_etypes]

    @property
    def is_homogeneous(self) -> bool:
        return len(self.ntypes) == 1 and len(self.canonical_etypes) == 1

    @property
    def independent(self) -> bool:
        """True when built by :func:`pack` (per-graph relation sets / readout masks)."""
        return self._rel_present is not None

    def num_nodes(self, ntype: Optional[str] = None) -> int:
        if ntype is None:
            return sum(self._num_nodes.values())
        return self._num_nodes[ntype]

    number_of_nodes = num_nodes

    def num_edges(self, etype=None) -> int:
        if etype is None:
            return sum(int(s.shape[0]) for s, _ in self._edges.values())
        return int(self._edges[self._canon(etype)][0].shape[0])

    number_of_edges = num_edges

    def _canon(self, etype) -> CEType:
        if isinstance(etype, tuple):
            if etype not in self._edges:
                raise KeyError(etype)
            return etype
        hits = [ce for ce in self.canonical_etypes if ce[1] == etype]
        if len(hits) != 1:
            raise KeyError(f"edge type {etype!r} is ambiguous or unknown; use the canonical triple")
        return hits[0]

    def edges(self, etype=None):
        if etype is None:
            if len(self.canonical_etypes) != 1:
                raise ValueError("etype is required for a graph with several relations")
            etype = self.canonical_etypes[0]
        return self._edges[self._canon(etype)]

    def __getitem__(self, key) -> RelationView:
        return RelationView(self, self._canon(key))

    @property
    def nodes(self) -> _NodesAccessor:
        return _NodesAccessor(self)

    @property
    def ndata(self) -> _TypedDataView:
        return _TypedDataView(self._ndata, self.ntypes)

    @property
    def edata(self) -> _TypedDataView:
        return _TypedDataView(self._edata, self.canonical_etypes)

    @property
    def device(self) -> torch.device:
        for fr in self._ndata.values():
            for v in fr.values():
                return v.device
        for s, _ in self._edges.values():
            return s.device
        return torch.device("cpu")

    def batch_num_nodes(self, ntype: Optional[str] = None) -> torch.Tensor:
        if ntype is None:
            if len(self.ntypes) != 1:
                raise ValueError("ntype is required")
            ntype = self.ntypes[0]
        return torch.tensor(self._batch_num_nodes[ntype], dtype=torch.int64)

    def batch_num_edges(self, etype=None) -> torch.Tensor:
        ce = self._canon(etype) if etype is not None else self.canonical_etypes[0]
        return torch.tensor(self._batch_num_edges[ce], dtype=torch.int64)

    @contextlib.contextmanager
    def local_scope(self):
        """Feature writes made inside the scope are dropped on exit ([DGL-mem] DGLGraph.local_scope)."""
        nsnap = {k: dict(v) for k, v in self._ndata.items()}
        esnap = {k: dict(v) for k, v in self._edata.items()}
        try:
            yield self
        finally:
            for k in self._ndata:
                self._ndata[k].clear()
                self._ndata[k].update(nsnap[k])
            for k in self._edata:
                self._edata[k].clear()
                self._edata[k].update(esnap[k])

    # ------------------------------------------------------------------ movement / IO
    def to(self, device, non_blocking: bool = False) -> "HeteroGraph":
        device = torch.device(device)
        if device == self.device and self._plan is not None and self._plan.device == device:
            return self
        g = HeteroGraph.__new__(HeteroGraph)
        g.ntypes = list(self.ntypes)
        g._num_nodes = dict(self._num_nodes)
        g.canonical_etypes = list(self.canonical_etypes)
        g._edges = {ce: (s.to(device, non_blocking=non_blocking), d.to(device, non_blocking=non_blocking))
                    for ce, (s, d) in self._edges.items()}
        g._ndata = {nt: _Frame({k: v.to(device, non_blocking=non_blocking) for k, v in fr.items()})
                    for nt, fr in self._ndata.items()}
        g._edata = {ce: _Frame({k: v.to(device, non_blocking=non_blocking) for k, v in fr.items()})
                    for ce, fr in self._edata.items()}
        g._batch_num_nodes = {k: list(v) for k, v in self._batch_num_nodes.items()}
        g._batch_num_edges = {k: list(v) for k, v in self._batch_num_edges.items()}
        g.batch_size = self.batch_size
        g._rel_present = None if self._rel_present is None else [list(r) for r in self._rel_present]
        g._plan = None
        return g

    def cpu(self):
        return self.to("cpu")

    def cuda(self):
        return self.to("cuda")

    def state(self) -> Dict:
        """A plain-dict form (tensors + python scalars) that ``torch.save`` can write."""
        return {
            "format": "wsi_hgnn_b200.HeteroGraph/1",
            "num_nodes": dict(self._num_nodes),
            "edges": {"|".join(ce): (s.cpu(), d.cpu()) for ce, (s, d) in self._edges.items()},
            "ndata": {nt: {k: v.cpu() for k, v in fr.items()} for nt, fr in self._ndata.items()},
            "edata": {"|".join(ce): {k: v.cpu() for k, v in fr.items()} for ce, fr in self._edata.items()},
            "batch_num_nodes": self._batch_num_nodes,
            "batch_num_edges": {"|".join(ce): v for ce, v in self._batch_num_edges.items()},
            "batch_size": self.batch_size,
            "rel_present": self._rel_present,
        }

    @staticmethod
    def from_state(st: Dict) -> "HeteroGraph":
        edges = {tuple(k.split("|")): v for k, v in st["edges"].items()}
        g = HeteroGraph(st["num_nodes"], edges, st["ndata"],
                        {tuple(k.split("|")): v for k, v in st["edata"].items()})
        g._batch_num_nodes = {k: list(v) for k, v in st["batch_num_nodes"].items()}
        g._batch_num_edges = {tuple(k.split("|")): list(v) for k, v in st["batch_num_edges"].items()}
        g.batch_size = st["batch_size"]
        g._rel_present = st.get("rel_present")
        return g

    def save(self, path: str):
        torch.save(self.state(), path)

    @staticmethod
    def load(path: str) -> "HeteroGraph":
        return HeteroGraph.from_state(torch.load(path, weights_only=False))

    @staticmethod
    def from_dgl(g) -> "HeteroGraph":
        """Duck-typed conversion of a DGLHeteroGraph (only usable where DGL exists; data.py:96-97)."""
        num_nodes = {nt: int(g.num_nodes(nt)) for nt in g.ntypes}
        edges, edata = {}, {}
        for ce in g.canonical_etypes:
            s, d = g.edges(etype=ce)
            edges[tuple(ce)] = (s, d)
            edata[tuple(ce)] = {k: v for k, v in g.edges[ce].data.items()}
        ndata = {nt: {k: v for k, v in g.nodes[nt].data.items()} for nt in g.ntypes}
        out = HeteroGraph(num_nodes, edges, ndata, edata)
        try:
            bs = int(g.batch_size)
            if bs > 1:
                out.batch_size = bs
                out._batch_num_nodes = {nt: [int(x) for x in g.batch_num_nodes(nt)] for nt in g.ntypes}
                out._batch_num_edges = {tuple(ce): [int(x) for x in g.batch_num_edges(ce)]
                                        for ce in g.canonical_etypes}
        except Exception:
            pass
        return out

    # ------------------------------------------------------------------ packed views
    def type_ptr(self) -> List[int]:
        ptr = [0]
        for nt in self.ntypes:
            ptr.append(ptr[-1] + self._num_nodes[nt])
        return ptr

    def packed_ndata(self, name: str = "feat", dtype=torch.float32) -> torch.Tensor:
        """[N, F] type-major packed copy of a node feature (rows of empty types are simply absent)."""
        parts = [self._ndata[nt][name] for nt in self.ntypes if self._num_nodes[nt] > 0]
        if not parts:
            raise ValueError("graph has no nodes")
        whole = getattr(self, "_packed_feat_view", None)
        if whole is not None and whole.dtype == dtype and whole.is_contiguous() and whole.shape[0] == sum(p.shape[0] for p in parts):
            # flat-format graphs (slide_io.FlatSlide): the per-type tensors are consecutive row slices of one packed
            # buffer; hand that buffer out instead of concatenating a copy (checked by address, not assumed)
            off, ok = whole.data_ptr(), True
            for p in parts:
                ok = ok and p.dtype == dtype and p.is_contiguous() and p.data_ptr() == off and p.shape[1:] == whole.shape[1:]
                off += p.numel() * p.element_size()
            if ok:
                return whole
        out = torch.cat([p.to(dtype) for p in parts], 0)
        return out.contiguous()

    def invalidate_plan(self):
        self._plan = None

    def _edata_signature(self, name: str):
        flat = getattr(self, "_flat_edges", None)
        if flat is not None:                            # flat-format graph: one tensor behind every relation's view
            return (flat[2].data_ptr(), flat[2]._version, len(self._edata))
        return tuple((fr[name].data_ptr(), fr[name]._version) if name in fr else None for fr in self._edata.values())

    def ndata_signature(self, name: str):
        """(data_ptr, version) of every node type's `name` tensor: changes when a feature is replaced or written in place."""
        return tuple((fr[name].data_ptr(), fr[name]._version) if name in fr else None for fr in self._ndata.values())

    def plan(self, sim_name: str = "sim") -> GraphPlan:
        """Build (once) the device layout the kernels consume; see :class:`GraphPlan`.

        Replaces what ``dgl.to_heterogeneous`` + DGL's on-demand CSC conversion do for the
        reference (construct_graph/graph_constructor.py:285-297; [DGL-mem] HeteroGraph::GetCSCMatrix).
        """
        if self._plan is not None:
            # the plan bakes the edge attribute into the CSR: a write to G.edata[sim_name] since then (a new tensor or an
            # in-place update - both visible as (data_ptr, _version)) makes it stale; the reference re-reads
            # G.edata['sim'] on every forward (models/HEATNet4.py:209-210)
            if getattr(self._plan, "_sim_sig", None) == self._edata_signature(sim_name):
                return self._plan
            self._plan = None
        dev = self.device
        p = GraphPlan()
        p.device = dev
        p._sim_sig = self._edata_signature(sim_name)
        p.ntypes = list(self.ntypes)
        p.rel_list = list(self.canonical_etypes)
        if len(p.rel_list) > 255:
            raise ValueError("at most 255 relations are supported")
        p.type_ptr = self.type_ptr()
        p.N = p.type_ptr[-1]
        p.B = self.batch_size
        tix = {nt: i for i, nt in enumerate(self.ntypes)}
        p.rel_src_type = [tix[ce[0]] for ce in p.rel_list]
        p.rel_dst_type = [tix[ce[2]] for ce in p.rel_list]
        T = len(self.ntypes)
        p.r_count = [sum(1 for d in p.rel_dst_type if d == t) for t in range(T)]

        # (type, graph) readout segments
        seg = [0]
        nonempty = torch.zeros(T, p.B, dtype=torch.bool)
        for t, nt in enumerate(self.ntypes):
            for b, n in enumerate(self._batch_num_nodes[nt]):
                seg.append(seg[-1] + n)
                nonempty[t, b] = n > 0
        assert seg[-1] == p.N
        p.seg_ptr_host = seg
        p.seg_nonempty = nonempty
        # 1/R per (type, graph) segment.  DGL batch: R_t of the shared metagraph.  pack(): per graph.
        seg_inv = []
        for t, nt in enumerate(self.ntypes):
            for b, n in enumerate(self._batch_num_nodes[nt]):
                if self._rel_present is None:
                    r = p.r_count[t]
                else:
                    r = sum(1 for ri, d in enumerate(p.rel_dst_type) if d == t and self._rel_present[b][ri])
                seg_inv.append(1.0 / r if r > 0 else 0.0)
        # every small host-side array of the plan travels in ONE buffer (pinned + non-blocking on CUDA: a pageable
        # torch.tensor(..., device=cuda) synchronises the stream, which would serialise the streaming evaluator)
        n_seg = len(seg) - 1
        R = len(p.rel_list)
        eptr = [0]
        for ce in p.rel_list:
            eptr.append(eptr[-1] + int(self._edges[ce][0].numel()))
        table = eptr + [p.type_ptr[t] for t in p.rel_src_type] + [0] + [p.type_ptr[t] for t in p.rel_dst_type] + [0]
        n0, n1, n2 = len(seg), len(seg) + len(p.type_ptr), len(seg) + len(p.type_ptr) + n_seg
        head = torch.empty(n2 + len(table), dtype=torch.int32)
        head[:n0] = torch.tensor(seg, dtype=torch.int32)
        head[n0:n1] = torch.tensor(p.type_ptr, dtype=torch.int32)
        head[n1:n2] = torch.tensor(seg_inv, dtype=torch.float32).view(torch.int32)
        head[n2:] = torch.tensor(table, dtype=torch.int32)
        head = _to_device_async(head, dev)
        p.seg_ptr = head[:n0]
        p.type_ptr_dev = head[n0:n1]
        seg_inv_dev = head[n1:n2].view(torch.float32)
        p._rel_table = head[n2:].view(3, R + 1) if (n2 * 4) % 16 == 0 else head[n2:].clone().view(3, R + 1)
        if p.N > 0:
            lens = (p.seg_ptr[1:] - p.seg_ptr[:-1]).to(torch.int64)
            p.node_inv_r = torch.repeat_interleave(seg_inv_dev, lens, output_size=p.N).contiguous()
        else:
            p.node_inv_r = torch.zeros(0, dtype=torch.float32, device=dev)

        # dst-major, relation-grouped CSR
        p.E = sum(int(self._edges[ce][0].numel()) for ce in p.rel_list)
        if dev.type == "cuda":
            self._plan_csr_native(p, sim_name)
        else:
            self._plan_csr_host(p, sim_name)
        self._plan = p
        return p

    def _plan_csr_native(self, p: GraphPlan, sim_name: str):
        """CSR via the plan-builder kernels (wsi_plan_build_csr): a handful of launches on the end-to-end path."""
        from . import ops
        dev = p.device
        R = len(p.rel_list)
        if p.E == 0:
            p.e_src = torch.zeros(0, dtype=torch.int32, device=dev)
            p.e_sim = torch.zeros(0, dtype=torch.float32, device=dev)
            p.e_rel = torch.zeros(0, dtype=torch.uint8, device=dev)
            p.rowptr = torch.zeros(p.N + 1, dtype=torch.int32, device=dev)
            return
        srcs = [self._edges[ce][0] for ce in p.rel_list]
        dsts = [self._edges[ce][1] for ce in p.rel_list]
        have_sim = [sim_name in self._edata[ce] for ce in p.rel_list]
        sim = None
        flat = self._flat_edge_views(p, srcs, dsts, sim_name, have_sim)
        if flat is not None:                  # flat-format graph: the per-relation tensors are slices of three arrays
            src, dst, sim = flat
            p.rowptr, p.e_src, p.e_sim, p.e_rel, _, stats = ops.plan_build_csr(src, dst, sim, p._rel_table, R, p.N)
            p._stats = stats
            return
        if all(have_sim):
            sims = [self._edata[ce][sim_name].reshape(-1) for ce in p.rel_list]
            if any(s.dtype != sims[0].dtype for s in sims) or sims[0].dtype not in (torch.float32, torch.float64):
                sims = [s.to(torch.float32) for s in sims]
            sim = torch.cat(sims) if R > 1 else sims[0].contiguous()
        elif any(have_sim):
            sim = torch.cat([self._edata[ce][sim_name].reshape(-1).to(torch.float32) if h else
                             torch.zeros(self._edges[ce][0].numel(), dtype=torch.float32, device=dev)
                             for ce, h in zip(p.rel_list, have_sim)])
        src = (torch.cat(srcs) if R > 1 else srcs[0]).to(torch.int64).contiguous()
        dst = (torch.cat(dsts) if R > 1 else dsts[0]).to(torch.int64).contiguous()
        p.rowptr, p.e_src, p.e_sim, p.e_rel, _, stats = ops.plan_build_csr(src, dst, sim, p._rel_table, R, p.N)
        p._stats = stats                  # [0] max in-degree, [1] range-error flag: read lazily (no sync here)

    def _flat_edge_views(self, p: GraphPlan, srcs, dsts, sim_name: str, have_sim):
        """(src, dst, sim) of ALL relations without concatenating, when the per-relation tensors are consecutive
        slices of the three arrays a FlatSlide graph was built on (checked by address, not assumed)."""
        flat = getattr(self, "_flat_edges", None)
        if flat is None or not all(have_sim):
            return None
        src, dst, sim = flat
        if src.numel() != p.E or src.dtype != torch.int64 or sim.dtype != torch.float32:
            return None
        o_s, o_d, o_m = src.data_ptr(), dst.data_ptr(), sim.data_ptr()
        for ce, s, d in zip(p.rel_list, srcs, dsts):
            m = self._edata[ce][sim_name]
            n = s.numel()
            if n and (s.data_ptr() != o_s or d.data_ptr() != o_d or m.data_ptr() != o_m or m.dtype != torch.float32
                      or s.dtype != torch.int64 or d.dtype != torch.int64 or m.numel() != n):
                return None
            o_s += 8 * n
            o_d += 8 * n
            o_m += 4 * n
        return src, dst, sim

    def _plan_csr_host(self, p: GraphPlan, sim_name: str):
        """Host-side (torch ops) statement of the same layout: CPU graphs in the tests, and the cross-check of the
        native builder."""
        dev = p.device
        srcs, dsts, rels, sims = [], [], [], []
        for ri, ce in enumerate(p.rel_list):
            s, d = self._edges[ce]
            if s.numel() == 0:
                continue
            srcs.append(s.to(dev) + p.type_ptr[p.rel_src_type[ri]])
            dsts.append(d.to(dev) + p.type_ptr[p.rel_dst_type[ri]])
            rels.append(torch.full((s.numel(),), ri, dtype=torch.int64, device=dev))
            if sim_name in self._edata[ce]:
                sims.append(self._edata[ce][sim_name].to(dev).reshape(-1).to(torch.float32))
            else:
                sims.append(torch.zeros(s.numel(), dtype=torch.float32, device=dev))
        if srcs:
            src = torch.cat(srcs)
            dst = torch.cat(dsts)
            rel = torch.cat(rels)
            sim = torch.cat(sims)
            if int(src.max()) >= p.N or int(dst.max()) >= p.N or int(src.min()) < 0 or int(dst.min()) < 0:
                raise IndexError("edge endpoint out of range")
            key = dst * 256 + rel                      # stable sort keeps original edge order inside a segment
            order = torch.argsort(key, stable=True)
            p.e_src = src[order].to(torch.int32).contiguous()
            p.e_sim = sim[order].contiguous()
            p.e_rel = rel[order].to(torch.uint8).contiguous()
            counts = torch.bincount(dst, minlength=p.N)
            p.max_in_degree = int(counts.max()) if p.N > 0 else 0
            rowptr = torch.zeros(p.N + 1, dtype=torch.int32, device=dev)
            rowptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
            p.rowptr = rowptr
        else:
            p.e_src = torch.zeros(0, dtype=torch.int32, device=dev)
            p.e_sim = torch.zeros(0, dtype=torch.float32, device=dev)
            p.e_rel = torch.zeros(0, dtype=torch.uint8, device=dev)
            p.rowptr = torch.zeros(p.N + 1, dtype=torch.int32, device=dev)

    def __repr__(self):
        return (f"HeteroGraph(num_nodes={self._num_nodes}, num_edges="
                f"{ {ce: int(s.shape[0]) for ce, (s, _) in self._edges.items()} }, batch_size={self.batch_size})")


def heterograph(data_dict: Dict[CEType, Tuple[torch.Tensor, torch.Tensor]],
                num_nodes_dict: Optional[Dict[str, int]] = None) -> HeteroGraph:
    """``dgl.heterograph``-style constructor."""
    if num_nodes_dict is None:
        num_nodes_dict = {}
        for (s, _, d), (u, v) in data_dict.items():
            u = torch.as_tensor(u)
            v = torch.as_tensor(v)
            num_nodes_dict[s] = max(num_nodes_dict.get(s, 0), int(u.max()) + 1 if u.numel() else 0)
            num_nodes_dict[d] = max(num_nodes_dict.get(d, 0), int(v.max()) + 1 if v.numel() else 0)
    return HeteroGraph(num_nodes_dict, data_dict)


def to_heterogeneous(src: torch.Tensor, dst: torch.Tensor, node_type: torch.Tensor, edge_type: torch.Tensor,
                     ntypes: Sequence[str], etypes: Sequence[str],
                     ndata: Optional[Dict[str, torch.Tensor]] = None,
                     edata: Optional[Dict[str, torch.Tensor]] = None) -> HeteroGraph:
    """Homogeneous (src, dst, _TYPE) -> HeteroGraph, as ``dgl.to_heterogeneous`` does for the
    reference builder (construct_graph/graph_constructor.py:285-297).

    [DGL-mem] all ``ntypes`` exist afterwards (possibly with 0 nodes); only the (s, e, d)
    triples that occur become relations; local ids are the rank of a node within its type
    (stable partition); ``_ID`` records the original ids.
    """
    node_type = torch.as_tensor(node_type, dtype=torch.int64)
    edge_type = torch.as_tensor(edge_type, dtype=torch.int64)
    src = torch.as_tensor(src, dtype=torch.int64)
    dst = torch.as_tensor(dst, dtype=torch.int64)
    N = node_type.numel()
    T = len(ntypes)
    order = torch.argsort(node_type, stable=True)
    counts = torch.bincount(node_type, minlength=T)
    starts = torch.cumsum(counts, 0) - counts
    local = torch.empty(N, dtype=torch.int64, device=node_type.device)
    local[order] = torch.arange(N, device=node_type.device) - starts[node_type[order]]
    num_nodes = {ntypes[t]: int(counts[t]) for t in range(T)}
    nfr = {}
    for t in range(T):
        ids = order[starts[t]:starts[t] + counts[t]]
        fr = {"_ID": ids}
        for k, v in (ndata or {}).items():
            fr[k] = v[ids]
        nfr[ntypes[t]] = fr
    st, dt = node_type[src], node_type[dst]
    key = (st * len(etypes) + edge_type) * T + dt
    edges, efr = {}, {}
    for kv in torch.unique(key).tolist():
        m = torch.nonzero(key == kv).reshape(-1)
        s_t = kv // (len(etypes) * T)
        e_t = (kv // T) % len(etypes)
        d_t = kv % T
        ce = (ntypes[s_t], etypes[e_t], ntypes[d_t])
        edges[ce] = (local[src[m]], local[dst[m]])
        fr = {"_ID": m}
        for k, v in (edata or {}).items():
            fr[k] = v[m]
        efr[ce] = fr
    return HeteroGraph(num_nodes, edges, nfr, efr)


def _cat_graphs(graphs: Sequence[HeteroGraph], rel_union: List[CEType]) -> HeteroGraph:
    ntypes = graphs[0].ntypes
    for g in graphs:
        if g.ntypes != ntypes:
            raise ValueError("all graphs must have the same node types")
    num_nodes = {nt: sum(g._num_nodes[nt] for g in graphs) for nt in ntypes}
    offs = {nt: [0] for nt in ntypes}
    for g in graphs:
        for nt in ntypes:
            offs[nt].append(offs[nt][-1] + g._num_nodes[nt])
    edges = {}
    edata: Dict[CEType, Dict[str, torch.Tensor]] = {}
    bne = {}
    dev = graphs[0].device
    for ce in rel_union:
        ss, dd, cnt = [], [], []
        keys = None
        for gi, g in enumerate(graphs):
            if ce in g._edges:
                s, d = g._edges[ce]
                ss.append(s + offs[ce[0]][gi])
                dd.append(d + offs[ce[2]][gi])
                cnt.append(int(s.shape[0]))
                if keys is None:
                    keys = set(g._edata[ce].keys())
                else:
                    keys &= set(g._edata[ce].keys())
            else:
                cnt.append(0)
        edges[ce] = (torch.cat(ss) if ss else torch.zeros(0, dtype=torch.int64, device=dev),
                     torch.cat(dd) if dd else torch.zeros(0, dtype=torch.int64, device=dev))
        bne[ce] = cnt
        edata[ce] = {}
        for k in sorted(keys or ()):
            edata[ce][k] = torch.cat([g._edata[ce][k] for g in graphs if ce in g._edges])
    ndata: Dict[str, Dict[str, torch.Tensor]] = {}
    for nt in ntypes:
        keys = None
        for g in graphs:
            ks = set(g._ndata[nt].keys())
            keys = ks if keys is None else keys & ks
        ndata[nt] = {k: torch.cat([g._ndata[nt][k] for g in graphs]) for k in sorted(keys or ())}
    out = HeteroGraph(num_nodes, edges, ndata, edata)
    out.batch_size = len(graphs)
    out._batch_num_nodes = {nt: [g._num_nodes[nt] for g in graphs] for nt in ntypes}
    out._batch_num_edges = bne
    return out


def batch(graphs: Sequence[HeteroGraph]) -> HeteroGraph:
    """``dgl.batch``: graphs must share node types AND the relation set ([DGL-mem])."""
    graphs = list(graphs)
    if not graphs:
        raise ValueError("empty batch")
    for g in graphs:
        if g.batch_size != 1:
            raise ValueError("batching already-batched graphs is not supported")
        if g.canonical_etypes != graphs[0].canonical_etypes:
            raise ValueError("dgl.batch semantics: all graphs must have the same relations; use pack()")
    return _cat_graphs(graphs, list(graphs[0].canonical_etypes))


def pack(graphs: Sequence[HeteroGraph]) -> HeteroGraph:
    """Block-diagonal packing whose forward equals ``torch.cat([gnn(g) for g in graphs])``
    (the reference trainer's tuple branch, trainer/train_gnn.py:59-62)."""
    graphs = list(graphs)
    if not graphs:
        raise ValueError("empty pack")
    for g in graphs:
        if g.batch_size != 1:
            raise ValueError("packing already-batched graphs is not supported")
    union = sorted(set(ce for g in graphs for ce in g.canonical_etypes))
    out = _cat_graphs(graphs, union)
    out._rel_present = [[ce in g._edges for ce in union] for g in graphs]
    return out


def unbatch(G: HeteroGraph) -> List[HeteroGraph]:
    """Inverse of :func:`batch` / :func:`pack` (``dgl.unbatch``): the per-slide graphs, with the relation
    set each one had (pack) or the shared one (batch)."""
    B = G.batch_size
    if B == 1:
        return [G]
    noff = {nt: [0] for nt in G.ntypes}
    for nt in G.ntypes:
        for n in G._batch_num_nodes[nt]:
            noff[nt].append(noff[nt][-1] + n)
    eoff = {ce: [0] for ce in G.canonical_etypes}
    for ce in G.canonical_etypes:
        for n in G._batch_num_edges[ce]:
            eoff[ce].append(eoff[ce][-1] + n)
    out = []
    for b in range(B):
        num_nodes = {nt: G._batch_num_nodes[nt][b] for nt in G.ntypes}
        edges, edata = {}, {}
        for ri, ce in enumerate(G.canonical_etypes):
            if G._rel_present is not None and not G._rel_present[b][ri]:
                continue
            a, z = eoff[ce][b], eoff[ce][b + 1]
            s, d = G._edges[ce]
            edges[ce] = (s[a:z] - noff[ce[0]][b], d[a:z] - noff[ce[2]][b])
            edata[ce] = {k: v[a:z] for k, v in G._edata[ce].items()}
        ndata = {nt: {k: v[noff[nt][b]:noff[nt][b + 1]] for k, v in G._ndata[nt].items()} for nt in G.ntypes}
        out.append(HeteroGraph(num_nodes, edges, ndata, edata))
    return out
