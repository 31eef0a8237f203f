"""Flat slide format + streaming inference over host-resident slides.

The reference stores every slide as a pickled DGLHeteroGraph (get_graph.py:279-289; read back by data.py:96-97) and
evaluates one graph at a time with a synchronous `.to(device)` + D2H read per slide (evaluator/eval_homo_graph.py:
61-95).  Here a slide is ONE contiguous byte blob + a small JSON-able header:

    [ features  fp32 or fp16 [N, F]  type-major packed (the GraphPlan row order) ]
    [ src       int64 [E]    local ids, relation-major (canonical_etypes order) ]
    [ dst       int64 [E] ]
    [ sim       fp32  [E] ]

so that (a) a file is written / read with one sequential IO (or mmap'ed), (b) the host -> device move is ONE
cudaMemcpyAsync from pinned memory instead of one per tensor (a config-2 slide has ~60 tensors), and (c) the device
HeteroGraph is a set of zero-copy views of the device blob.  `stream_forward` runs the model over an iterable of
flat slides with the copy of slide i+1 (copy engine, its own stream) overlapping the forward of slide i, and the
logits of slide i-1 read back asynchronously: the per-epoch evaluation loop without a host sync per slide.
"""
import json
import os
import queue
import struct
import threading
from typing import Dict, Iterable, Iterator, List, Optional, Tuple

import numpy as np
import torch

from .hetero_graph import HeteroGraph

MAGIC = b"WSIFLAT1"
_ALIGN = 256


def _align(n: int) -> int:
    return (n + _ALIGN - 1) // _ALIGN * _ALIGN


class FlatSlide:
    """One slide in the flat format (host memory: pageable, pinned or an mmap of a file)."""

    def __init__(self, header: Dict, blob: torch.Tensor):
        if blob.dtype != torch.uint8 or blob.dim() != 1:
            raise TypeError("FlatSlide: blob must be a 1-D uint8 tensor")
        if blob.numel() < header["nbytes"]:
            raise ValueError("FlatSlide: blob shorter than the header says")
        self.header, self.blob = header, blob

    # ------------------------------------------------------------------ construction
    @staticmethod
    def from_graph(G: HeteroGraph, feat_name: str = "feat", sim_name: str = "sim", pin: bool = False,
                   feat_dtype: str = "fp32") -> "FlatSlide":
        """feat_dtype "fp16": the features are stored as fp16 (round to nearest, clamped to +-65504) - exactly the operand
        the default single-pass fp16 GEMM forms from fp32 features, so the logits are bit-identical while the blob (and
        the host -> device copy that bounds the streamed path) is half the size; the tensor-core chain then reads the
        features in place (no conversion pass)."""
        if feat_dtype not in ("fp32", "fp16"):
            raise ValueError("feat_dtype must be 'fp32' or 'fp16'")
        fsz = 4 if feat_dtype == "fp32" else 2
        fdt = torch.float32 if feat_dtype == "fp32" else torch.float16
        if G.batch_size != 1:
            raise ValueError("FlatSlide holds one slide; flatten the graphs before batching / packing them")
        ntypes, cets = list(G.ntypes), list(G.canonical_etypes)
        n_per = [G.num_nodes(nt) for nt in ntypes]
        N = sum(n_per)
        feats = [G.nodes[nt].data[feat_name] for nt in ntypes if G.num_nodes(nt) > 0]
        F = int(feats[0].shape[1]) if feats else 0
        e_per = [int(G._edges[ce][0].numel()) for ce in cets]
        E = sum(e_per)
        off_feat = 0
        off_src = _align(off_feat + N * F * fsz)
        off_dst = _align(off_src + E * 8)
        off_sim = _align(off_dst + E * 8)
        nbytes = _align(off_sim + E * 4)
        blob = torch.zeros(nbytes, dtype=torch.uint8)
        if pin and torch.cuda.is_available():
            blob = blob.pin_memory()
        if N * F:
            allf = torch.cat([f.detach().to("cpu", torch.float32) for f in feats], 0)
            if fsz == 2:
                allf = allf.clamp(-65504.0, 65504.0)
            blob[off_feat:off_feat + N * F * fsz].view(fdt).view(N, F).copy_(allf)
        if E:
            src = torch.cat([G._edges[ce][0].detach().to("cpu", torch.int64) for ce in cets])
            dst = torch.cat([G._edges[ce][1].detach().to("cpu", torch.int64) for ce in cets])
            sim = torch.cat([(G._edata[ce][sim_name].detach().reshape(-1).to("cpu", torch.float32) if sim_name in G._edata[ce]
                              else torch.zeros(n, dtype=torch.float32)) for ce, n in zip(cets, e_per)])
            blob[off_src:off_src + E * 8].view(torch.int64).copy_(src)
            blob[off_dst:off_dst + E * 8].view(torch.int64).copy_(dst)
            blob[off_sim:off_sim + E * 4].view(torch.float32).copy_(sim)
        header = {"format": "wsi_hgnn_b200.FlatSlide/1", "ntypes": ntypes, "num_nodes": n_per,
                  "canonical_etypes": [list(ce) for ce in cets], "num_edges": e_per, "feat_dim": F,
                  "feat_name": feat_name, "sim_name": sim_name, "feat_dtype": feat_dtype,
                  "off": {"feat": off_feat, "src": off_src, "dst": off_dst, "sim": off_sim}, "nbytes": nbytes}
        return FlatSlide(header, blob)

    def pin(self) -> "FlatSlide":
        return self if self.blob.is_pinned() else FlatSlide(self.header, self.blob.pin_memory())

    # ------------------------------------------------------------------ file IO
    def save(self, path: str):
        hb = json.dumps(self.header).encode()
        pad = _align(len(MAGIC) + 8 + len(hb)) - (len(MAGIC) + 8 + len(hb))
        with open(path, "wb") as f:
            f.write(MAGIC)
            f.write(struct.pack("<Q", len(hb) + pad))
            f.write(hb + b" " * pad)
            f.write(self.blob[:self.header["nbytes"]].numpy().tobytes())

    @staticmethod
    def load(path: str, mmap: bool = True) -> "FlatSlide":
        with open(path, "rb") as f:
            if f.read(len(MAGIC)) != MAGIC:
                raise ValueError(f"{path}: not a FlatSlide file")
            (hlen,) = struct.unpack("<Q", f.read(8))
            header = json.loads(f.read(hlen).decode())
            off = len(MAGIC) + 8 + hlen
            if mmap:
                arr = np.memmap(path, dtype=np.uint8, mode="r", offset=off, shape=(header["nbytes"],))
                blob = torch.from_numpy(np.asarray(arr))         # read-only view of the page cache
            else:
                blob = torch.frombuffer(bytearray(f.read(header["nbytes"])), dtype=torch.uint8)
        return FlatSlide(header, blob)

    # ------------------------------------------------------------------ views
    def graph_on(self, blob: torch.Tensor) -> HeteroGraph:
        """HeteroGraph whose tensors are zero-copy views of `blob` (this slide's bytes on any device)."""
        h = self.header
        off, F = h["off"], h["feat_dim"]
        n_per, e_per = h["num_nodes"], h["num_edges"]
        N, E = sum(n_per), sum(e_per)
        feat = self._feat_view(blob)
        src = blob[off["src"]:off["src"] + E * 8].view(torch.int64)
        dst = blob[off["dst"]:off["dst"] + E * 8].view(torch.int64)
        sim = blob[off["sim"]:off["sim"] + E * 4].view(torch.float32)
        ndata, edges, edata = {}, {}, {}
        r = 0
        for nt, n in zip(h["ntypes"], n_per):
            ndata[nt] = {h["feat_name"]: feat[r:r + n]}
            r += n
        e = 0
        for ce, n in zip(h["canonical_etypes"], e_per):
            edges[tuple(ce)] = (src[e:e + n], dst[e:e + n])
            edata[tuple(ce)] = {h["sim_name"]: sim[e:e + n]}
            e += n
        G = HeteroGraph(dict(zip(h["ntypes"], n_per)), edges, ndata, edata)
        G._packed_feat_view = feat                                # packed_ndata() returns it without a copy
        G._flat_edges = (src, dst, sim)                           # plan() reads them without concatenating
        return G

    def feat_is_fp16(self) -> bool:
        return self.header.get("feat_dtype", "fp32") == "fp16"

    def _feat_view(self, blob: torch.Tensor) -> torch.Tensor:
        h = self.header
        N, F, o = sum(h["num_nodes"]), h["feat_dim"], h["off"]["feat"]
        if self.feat_is_fp16():
            return blob[o:o + N * F * 2].view(torch.float16).view(N, F)
        return blob[o:o + N * F * 4].view(torch.float32).view(N, F)

    def _plan_head(self):
        """Host side of the plan of this (single) slide, computed once per FlatSlide: everything GraphPlan needs that
        follows from the header alone, and ONE pinned int32 buffer [seg_ptr | rel table | node_inv_r] for the device."""
        if getattr(self, "_head", None) is None:
            h = self.header
            ntypes, n_per = h["ntypes"], h["num_nodes"]
            rel_list = [tuple(ce) for ce in h["canonical_etypes"]]
            if len(rel_list) > 255:
                raise ValueError("at most 255 relations are supported")
            tix = {nt: i for i, nt in enumerate(ntypes)}
            type_ptr = [0]
            for n in n_per:
                type_ptr.append(type_ptr[-1] + n)
            src_t, dst_t = [tix[ce[0]] for ce in rel_list], [tix[ce[2]] for ce in rel_list]
            T, R, N = len(ntypes), len(rel_list), type_ptr[-1]
            r_count = [sum(1 for d in dst_t if d == t) for t in range(T)]
            eptr = [0]
            for n in h["num_edges"]:
                eptr.append(eptr[-1] + n)
            table = eptr + [type_ptr[t] for t in src_t] + [0] + [type_ptr[t] for t in dst_t] + [0]
            inv = np.repeat(np.array([1.0 / r if r > 0 else 0.0 for r in r_count], dtype=np.float32), n_per)
            n0 = T + 1
            n1 = n0 + len(table)
            n1p = (n1 + T + 3) // 4 * 4                              # node_inv_r starts 16 B aligned
            buf = torch.zeros(n1p + N, dtype=torch.int32)
            buf[:n0] = torch.tensor(type_ptr, dtype=torch.int32)
            buf[n0:n1] = torch.tensor(table, dtype=torch.int32)
            buf[n1:n1 + T] = torch.tensor([1.0 if n > 0 else 0.0 for n in n_per], dtype=torch.float32).view(torch.int32)
            if N:
                buf[n1p:] = torch.from_numpy(inv).view(torch.int32)
            if torch.cuda.is_available():
                buf = buf.pin_memory()
            self._head = dict(ntypes=list(ntypes), rel_list=rel_list, type_ptr=type_ptr, src_t=src_t, dst_t=dst_t,
                              r_count=r_count, N=N, E=eptr[-1], R=R, T=T, n0=n0, n1=n1, n1p=n1p, buf=buf,
                              nonempty=torch.tensor([[n > 0] for n in n_per], dtype=torch.bool).reshape(T, 1))
        return self._head

    def plan_on(self, blob: torch.Tensor):
        """(GraphPlan, packed features [N, F]) of this slide built straight from its device blob - the same plan
        HeteroGraph.plan() builds for `graph_on(blob)`, without materialising the per-type / per-relation views
        (~60 tensor objects for a config-2 slide): the host path of the streaming evaluator."""
        from . import ops
        from .hetero_graph import GraphPlan
        hd, h = self._plan_head(), self.header
        dev = blob.device
        if dev.type != "cuda":
            raise RuntimeError("FlatSlide.plan_on needs the blob on a CUDA device")
        off, F = h["off"], h["feat_dim"]
        N, E, R, T = hd["N"], hd["E"], hd["R"], hd["T"]
        head = hd["buf"].to(dev, non_blocking=True)
        p = GraphPlan()
        p.device, p.ntypes, p.rel_list = dev, list(hd["ntypes"]), list(hd["rel_list"])
        p.type_ptr, p.N, p.E, p.B = list(hd["type_ptr"]), N, E, 1
        p.rel_src_type, p.rel_dst_type, p.r_count = list(hd["src_t"]), list(hd["dst_t"]), list(hd["r_count"])
        p.seg_ptr_host, p.seg_nonempty = list(hd["type_ptr"]), hd["nonempty"]
        p.seg_ptr = p.type_ptr_dev = head[:hd["n0"]]
        p._rel_table = head[hd["n0"]:hd["n1"]].view(3, R + 1)
        p.node_inv_r = head[hd["n1p"]:].view(torch.float32)
        feat = self._feat_view(blob)
        if E == 0:
            p.e_src = torch.zeros(0, dtype=torch.int32, device=dev)
            p.e_sim = torch.zeros(0, dtype=torch.float32, device=dev)
            p.e_rel = torch.zeros(0, dtype=torch.uint8, device=dev)
            p.rowptr = torch.zeros(N + 1, dtype=torch.int32, device=dev)
            return p, feat
        src = blob[off["src"]:off["src"] + E * 8].view(torch.int64)
        dst = blob[off["dst"]:off["dst"] + E * 8].view(torch.int64)
        sim = blob[off["sim"]:off["sim"] + E * 4].view(torch.float32)
        p.rowptr, p.e_src, p.e_sim, p.e_rel, _, p._stats = ops.plan_build_csr(src, dst, sim, p._rel_table, R, N)
        return p, feat

    def to_graph(self, device="cpu", non_blocking: bool = False) -> HeteroGraph:
        dev = torch.device(device)
        blob = self.blob[:self.header["nbytes"]]
        return self.graph_on(blob if dev.type == "cpu" else blob.to(dev, non_blocking=non_blocking))

    def num_edges(self) -> int:
        return sum(self.header["num_edges"])

    def num_nodes(self) -> int:
        return sum(self.header["num_nodes"])


_STREAM_CTX: Dict = {}


def _stream_ctx(dev: torch.device, nbuf: int) -> Dict:
    """per-device streams + blob buffers of stream_forward, reused by successive (non-overlapping) calls."""
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), nbuf)
    ctx = _STREAM_CTX.get(key)
    if ctx is None or ctx["busy"]:
        ctx = {"copy": torch.cuda.Stream(device=dev), "plan": torch.cuda.Stream(device=dev), "bufs": [None] * nbuf,
               "busy": False}
        _STREAM_CTX.setdefault(key, ctx)
    ctx["busy"] = True
    return ctx


def stream_forward(model, slides: Iterable[FlatSlide], device, depth: int = 4, threaded: bool = False) -> Iterator[torch.Tensor]:
    """Yield the logits ([1, out_dim], host tensor) of every slide, in order.

    Three stages run concurrently on three streams, each ahead of the next:
        copy stream   ONE host -> device copy of the slide's blob (copy engine)
        plan stream   CSR + work-list build on the device blob (the one host read of the planner waits only for this
                      stream, i.e. for a copy that was queued a whole slide earlier)
        main stream   forward (wsi_heat_forward), logits -> pinned host memory (asynchronous)
    For HEATNet2 / HEATNet4 on shapes the tensor-core chain takes, the last two stages of a slide are issued by ONE C call
    (wsi_slide_forward); other models / shapes go through the Python-issued planner and forward (same kernels).
    The caller's thread interleaves the stages one slide apart; with `threaded` the first two stages are issued by a
    worker thread instead (planning is the larger part of the host cost per slide, but most of it holds the GIL: on
    config-2 slides it measured no faster - 0.97 vs 0.93 ms / slide - so it is not the default).  The host never waits for the main stream except on a
    slide's own `done` event, `depth` slides later.  `depth` = device blob buffers (>= 3).
    The blobs should be pinned (FlatSlide.pin()) for the copies to be asynchronous."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("stream_forward needs a CUDA device (there is no CPU fallback)")
    nbuf = max(4, int(depth))
    main = torch.cuda.current_stream(dev)
    # streams and device blob buffers persist across calls (per device): a new stream per call would make the caching
    # allocator cudaMalloc fresh buffers every time, and cudaMalloc stalls for tens of ms once CUDA graphs exist
    ctx = _stream_ctx(dev, nbuf)
    copy_stream, plan_stream, bufs = ctx["copy"], ctx["plan"], ctx["bufs"]
    was_training = model.training
    model.eval()

    def upload(k: int, s: FlatSlide, free_ev):
        n = s.header["nbytes"]
        with torch.cuda.stream(copy_stream):
            if bufs[k] is None or bufs[k].numel() < n:
                if free_ev is not None:
                    free_ev.synchronize()
                bufs[k] = torch.empty(int(n * 1.25) + _ALIGN, dtype=torch.uint8, device=dev)
            elif free_ev is not None:
                copy_stream.wait_event(free_ev)                     # the forward that last read this buffer
            bufs[k][:n].copy_(s.blob[:n], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return s, k, ev

    # plan + forward without a HeteroGraph per slide (WSI_STREAM_NO_FAST: development knob for A/B timing)
    fast = hasattr(model, "forward_planned") and not os.environ.get("WSI_STREAM_NO_FAST")

    def plan_begin(staged):
        """CSR build and the counting half of the work list: kernels only, no host wait"""
        s, k, uploaded = staged
        with torch.cuda.stream(plan_stream), torch.no_grad():
            plan_stream.wait_event(uploaded)
            blob = bufs[k][:s.header["nbytes"]]
            if fast and s.num_nodes() > 0:
                p, feat = s.plan_on(blob)
                G = (s, blob, p, feat)
            else:
                G = s.graph_on(blob)
                p = G.plan()
            if hasattr(model, "prepare_plan_begin"):
                model.prepare_plan_begin(p)
        return G, k, p

    def plan_finish(begun):
        """everything of the plan that needs a host read (by now the counting kernels had a forward's worth of head start)"""
        G, k, p = begun
        with torch.cuda.stream(plan_stream), torch.no_grad():
            if hasattr(model, "prepare_plan"):
                model.prepare_plan(p)
            ev = torch.cuda.Event()
            ev.record(plan_stream)
        return G, k, ev

    def plan(staged):
        return plan_finish(plan_begin(staged))

    def forward(planned):
        G, k, ready = planned
        main.wait_event(ready)
        with torch.no_grad():
            out = None
            if isinstance(G, tuple):
                s, blob, p, feat = G
                out = model.forward_planned(p, feat)
                if out is None:                                     # a shape / mode the one-call driver does not take
                    Gg = s.graph_on(blob)
                    Gg._plan = p
                    G = (Gg, blob, p, feat)
                    out = model(Gg)
            else:
                out = model(G)
        host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
        host.copy_(out, non_blocking=True)
        done = torch.cuda.Event()
        done.record(main)
        return host, done, G                                        # G (and its plan tensors, allocated on the plan
                                                                    # stream) stay alive until the forward has finished
    pending: List[Tuple[torch.Tensor, torch.cuda.Event, HeteroGraph]] = []
    stop = threading.Event()
    worker = None
    # blob -> logits in ONE host call per slide (wsi_slide_forward: planner + forward issued from C; measured 0.89-0.95
    # against 1.13-1.15 ms / slide for the Python-issued stages on the same box).  WSI_STREAM_NATIVE=0 (development knob)
    # or `threaded` select the Python-issued stages.
    native = (hasattr(model, "slide_plan_native") and os.environ.get("WSI_STREAM_NATIVE", "1") != "0" and not threaded)
    try:
        # the whole loop in ONE C call (wsi_stream_forward: copies, planner and forwards issued by native code, three
        # slides in flight) whenever every slide is the one-call driver's shape; WSI_STREAM_LOOP=python keeps the
        # Python-issued pipeline below (same kernels, same order)
        if native and hasattr(model, "stream_native") and os.environ.get("WSI_STREAM_LOOP", "native") == "native":
            it_all = iter(slides)
            while True:
                chunk = []
                for s_ in it_all:
                    chunk.append(s_)
                    if len(chunk) >= 256:
                        break
                if not chunk:
                    break
                with torch.no_grad():
                    outs = model.stream_native(chunk, dev, nbuf, ctx)
                if outs is None:                                    # not all slides fit: Python-issued pipeline for these
                    ctx["busy"] = False
                    old = os.environ.get("WSI_STREAM_LOOP")
                    os.environ["WSI_STREAM_LOOP"] = "python"
                    try:
                        outs = list(stream_forward(model, chunk, dev, depth, threaded))
                    finally:
                        if old is None:
                            os.environ.pop("WSI_STREAM_LOOP", None)
                        else:
                            os.environ["WSI_STREAM_LOOP"] = old
                    ctx["busy"] = True
                for o in outs:
                    yield o
        elif native:
            # software pipeline, one slide apart per stage, no host wait on anything younger than a whole slide:
            #   upload(i + 2)  ->  plan(i + 1): wsi_slide_plan (CSR + counting, totals -> pinned host, event)  ->
            #   run(i): wait for slide i's plan event (recorded one iteration ago), wsi_slide_run (fill + forward), logits D2H
            slots = ctx.setdefault("slots", [dict() for _ in range(nbuf)])
            free_ev: List[Optional[torch.cuda.Event]] = [None] * nbuf
            it = iter(slides)
            n_up = 0

            def next_upload():
                nonlocal n_up
                s_ = next(it, None)
                if s_ is None:
                    return None
                k_ = n_up % nbuf
                n_up += 1
                return upload(k_, s_, free_ev[k_])

            def plan_native(staged_):
                s_, k_, uploaded_ = staged_
                blob_ = bufs[k_][:s_.header["nbytes"]]
                state_ = head_ = None
                if s_.num_nodes() > 0:
                    with torch.cuda.stream(plan_stream), torch.no_grad():
                        plan_stream.wait_event(uploaded_)
                        head_ = s_._plan_head()["buf"].to(dev, non_blocking=True)
                        state_ = model.slide_plan_native(s_, blob_, head_, slots[k_], plan_stream)
                ev_ = torch.cuda.Event()
                ev_.record(plan_stream)
                return s_, k_, uploaded_, blob_, head_, state_, ev_

            up_q = [u for u in (next_upload(), next_upload()) if u is not None]
            cur = plan_native(up_q.pop(0)) if up_q else None
            while cur is not None:
                u = next_upload()                                   # stage 1: slide i + 2
                if u is not None:
                    up_q.append(u)
                nxt = plan_native(up_q.pop(0)) if up_q else None    # stage 2: slide i + 1 (kernels + async totals)
                s, k, uploaded, blob, head, state, planned_ev = cur
                with torch.no_grad():
                    if state is not None:
                        planned_ev.synchronize()                    # a whole slide old: the totals are on the host
                        out = model.slide_run_native(state, slots[k], plan_stream, main)
                        host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
                        host.copy_(out, non_blocking=True)
                        done = torch.cuda.Event()
                        done.record(main)
                        keep = (s, blob, head, state)
                    else:                                           # not the driver's shapes: generic path
                        host, done, keep = forward(plan((s, k, uploaded)))
                free_ev[k] = done
                pending.append((host, done, keep))
                if len(pending) >= nbuf:
                    h, ev, _ = pending.pop(0)
                    ev.synchronize()
                    yield h
                cur = nxt
        elif threaded:
            ready_q: "queue.Queue" = queue.Queue(maxsize=max(1, nbuf - 2))
            free_q = [queue.Queue() for _ in range(nbuf)]           # per buffer: `done` event of the forward that read it

            def produce():
                try:
                    torch.cuda.set_device(dev)
                    up = None
                    for i, s in enumerate(slides):
                        if stop.is_set():
                            return
                        k, free_ev = i % nbuf, None
                        if i >= nbuf:
                            while free_ev is None and not stop.is_set():
                                try:
                                    free_ev = free_q[k].get(timeout=0.05)
                                except queue.Empty:
                                    pass
                        nxt = upload(k, s, free_ev)                 # slide i's copy is queued before slide i-1 is planned
                        if up is not None:
                            ready_q.put(plan(up))
                        up = nxt
                    if up is not None and not stop.is_set():
                        ready_q.put(plan(up))
                    ready_q.put(None)
                except BaseException as e:                          # surfaces in the consumer
                    ready_q.put(e)

            worker = threading.Thread(target=produce, name="wsi-stream-plan", daemon=True)
            worker.start()
            while True:
                item = ready_q.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                host, done, G = forward(item)
                free_q[item[1]].put(done)
                pending.append((host, done, G))
                if len(pending) >= nbuf:
                    h, ev, _ = pending.pop(0)
                    ev.synchronize()
                    yield h
        else:
            free_ev: List[Optional[torch.cuda.Event]] = [None] * nbuf
            it = iter(slides)
            up: List = []                                           # uploaded, not yet planned
            n_up = 0
            for _ in range(2):
                s = next(it, None)
                if s is not None:
                    up.append(upload(n_up % nbuf, s, None))
                    n_up += 1
            planned = plan(up.pop(0)) if up else None
            while planned is not None:
                s = next(it, None)
                if s is not None:                                   # stage 1: slide i+2
                    up.append(upload(n_up % nbuf, s, free_ev[n_up % nbuf]))
                    n_up += 1
                begun = plan_begin(up.pop(0)) if up else None       # stage 2a: slide i+1, kernels only
                host, done, G = forward(planned)                    # stage 3: slide i
                nxt = plan_finish(begun) if begun is not None else None    # stage 2b: the host reads of slide i+1's plan
                free_ev[planned[1]] = done
                pending.append((host, done, G))
                if len(pending) >= nbuf:
                    h, ev, _ = pending.pop(0)
                    ev.synchronize()
                    yield h
                planned = nxt
        for h, ev, _ in pending:
            ev.synchronize()
            yield h
        pending = []
    finally:
        stop.set()
        if worker is not None:
            try:                                                    # unblock a producer waiting on a full queue
                while worker.is_alive():
                    try:
                        ready_q.get_nowait()
                    except queue.Empty:
                        worker.join(timeout=0.05)
            except Exception:
                pass
        if pending:                                                 # abandoned mid-way: let the queued forwards finish
            torch.cuda.current_stream(dev).synchronize()            # before their buffers are handed to the next call
        ctx["busy"] = False
        if was_training:
            model.train()
