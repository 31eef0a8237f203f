"""Data-parallel training over the GPUs of one box: slides are sharded over ranks (sharding.py), every rank runs
forward + backward on its own slides, and the ONE collective of the step is an all-reduce (sum) of the flat fp32
gradient buffer over NCCL / NVLink, issued bucket by bucket WHILE the backward is still running.
The reference has no distributed code at all (trainer/trainer.py:32-34 is single-device); this is the config-5
wrapper around its train_one_step (trainer/train_gnn.py:55-79).

Memory layout: `FlatModel` re-homes every trainable parameter into ONE flat fp32 buffer and gives it a gradient that is
a view of a second flat buffer, so that
  * autograd accumulates straight into the communication buffer (no pack / unpack launches),
  * zero_grad is one memset (or free: fused into the optimizer kernel),
  * Adam is ONE kernel of libwsi_hgnn.so over the four flat buffers (wsi_adam_step) instead of torch's
    multi_tensor_apply chain over ~100 tensors,
  * a bucket = a contiguous slice of the gradient buffer, all-reduced asynchronously as soon as the last gradient of
    the slice has been accumulated (post-accumulate-grad hooks).
"""
import os
from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


class FlatGradAllReduce:
    """Gradients of `params` as views of ONE flat buffer + a single all-reduce of it (the simple, non-overlapped form:
    CPU / gloo tests, and the fallback of FlatModel for parameters whose gradient arrives out of bucket order)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        self.flat: Optional[torch.Tensor] = None

    def _adopt(self):
        """(Re)point every p.grad at its slice of the flat buffer, keeping gradients that already exist."""
        dev = self.params[0].device
        if self.flat is None or self.flat.device != dev:
            self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            n = p.numel()
            view = self.flat[off:off + n].view_as(p)
            if p.grad is None:
                view.zero_()                            # unused in the forward (HEATLayer.weight, attn.*): zeros
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
            p.grad = view                               # every rank ends up with the reduced gradient, used or not
            off += n

    def __call__(self):
        if not self.params:
            return
        self._adopt()
        if _world(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)


class FlatModel:
    """Flat parameter / gradient / Adam-state buffers of `model` with bucketed, overlapped gradient all-reduce."""

    def __init__(self, model: torch.nn.Module, bucket_mb: float = 16.0, group=None):
        self.params = [p for p in model.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("FlatModel: the model has no trainable parameter")
        self.group = group
        dev = self.params[0].device
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4           # 16 B aligned slices
        self.numel = total
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.step_count = 0
        self.offs = offs
        self._active = [True] * len(self.params)        # parameters that received a gradient in the current step
        n = len(self.params)
        self.offs_dev = torch.tensor(offs + [total], dtype=torch.int64, device=dev)
        self.flags_dev = torch.ones(n, dtype=torch.int32, device=dev)       # _active on the device (MAX-reduced over the ranks)
        self.steps_dev = torch.zeros(n, dtype=torch.int32, device=dev)      # per-parameter Adam step (torch counts only steps with a gradient)
        self.corr_ws = torch.zeros(2 * n, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, offs):
                n = p.numel()
                self.flat_p[o:o + n].copy_(p.data.reshape(-1))
                p.data = self.flat_p[o:o + n].view_as(p)
        # gradient slices.  During a step p.grad is None, so autograd's AccumulateGrad adopts the incoming gradient (a view
        # of the stacked per-type weight gradient - no kernel) instead of launching one in-place add per parameter
        # (~150 launches per backward for HEATNet4); fold() then adds all of them into the flat buffer with one
        # multi-tensor launch per bucket.  After finish() p.grad is the flat slice (what the optimizer / a test reads).
        self.views = [self.flat_g[o:o + p.numel()].view_as(p) for p, o in zip(self.params, offs)]
        for p, v in zip(self.params, self.views):
            p.grad = v
        # buckets: contiguous slices, filled from the LAST parameter backwards (gradients arrive roughly in reverse
        # registration order), each >= bucket_mb
        limit = int(bucket_mb * (1 << 20) / 4)
        self.buckets = []                               # (start, end, [param indices])
        end, idx = total, []
        for i in range(len(self.params) - 1, -1, -1):
            idx.append(i)
            if end - offs[i] >= limit or i == 0:
                self.buckets.append((offs[i], end, list(idx)))
                end, idx = offs[i], []
        self.bucket_of = {}
        for b, (_, _, ids) in enumerate(self.buckets):
            for i in ids:
                self.bucket_of[i] = b
        self.expected = None                            # per bucket: parameters whose gradient arrived last step
        self._seen = [set() for _ in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._handles = []
        self._late = []
        self.armed = False                              # hooks fire collectives only in the last micro-batch of a step
        for i, p in enumerate(self.params):
            p.register_post_accumulate_grad_hook(self._make_hook(i))

    def _make_hook(self, i: int):
        def hook(_p):
            if not self.armed:
                return
            b = self.bucket_of[i]
            self._seen[b].add(i)
            if self._launched[b]:
                self._late.append(i)                    # arrived after its bucket was reduced (see finish)
            elif self.expected is not None and self._seen[b] >= self.expected[b] and _world(self.group) > 1:
                self._launch(b)
        return hook

    def begin_step(self):
        """Before the first backward of a step: detach the gradient slices so that autograd adopts instead of adds."""
        for p in self.params:
            p.grad = None
        self._active = [False] * len(self.params)

    def fold(self, ids=None):
        """flat_g += the gradients autograd left on the parameters `ids` (default: all), which are then detached again."""
        dst, src = [], []
        for i in (range(len(self.params)) if ids is None else ids):
            p = self.params[i]
            g = p.grad
            if g is None:                                                   # nothing arrived
                continue
            self._active[i] = True
            if g is self.views[i]:                                          # accumulated in place already
                continue
            dst.append(self.views[i])
            src.append(g if g.shape == self.views[i].shape else g.reshape(self.views[i].shape))
            p.grad = None
        if dst:
            with torch.no_grad():
                torch._foreach_add_(dst, src)

    def _launch(self, b: int):
        s, e, ids = self.buckets[b]
        self.fold(ids)
        self._launched[b] = True
        self._handles.append(dist.all_reduce(self.flat_g[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def arm(self):
        """Call before the backward of the step's LAST micro-batch."""
        self.armed = True
        self._seen = [set() for _ in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._handles = []
        self._late = []

    def finish(self):
        """After the last backward: reduce whatever the hooks did not (first step, parameters without a gradient), then
        make the compute stream wait for every bucket."""
        self.armed = False
        if _world(self.group) > 1:
            for b in range(len(self.buckets)):
                if not self._launched[b]:
                    self._launch(b)
            for h in self._handles:
                h.wait()
            # a gradient that arrived AFTER its bucket's early launch (a parameter that received none in the previous
            # step): its slice holds the reduced sum of nothing yet - add the local gradient and reduce that slice alone
            late = self._late
            if late:
                self.fold(late)
                for i in late:
                    dist.all_reduce(self.views[i], op=dist.ReduceOp.SUM, group=self.group)
        else:
            self.fold()
        self.expected = [set(s) for s in self._seen]
        self._handles = []
        # a parameter is stepped iff ANY rank produced a gradient for it (the single-process reference sees the whole
        # batch): the flags go to the device asynchronously and are MAX-reduced there - no host round trip
        h = torch.tensor([1 if a else 0 for a in self._active], dtype=torch.int32)
        if self.flat_g.is_cuda:
            h = h.pin_memory()            # (cached pinned block; the allocator keeps it until the copy below has run)
        self.flags_dev.copy_(h, non_blocking=True)
        if _world(self.group) > 1:
            dist.all_reduce(self.flags_dev, op=dist.ReduceOp.MAX, group=self.group)
        for p, v in zip(self.params, self.views):
            p.grad = v

    def zero_grad(self):
        self.flat_g.zero_()

    def adam_step(self, lr: float, weight_decay: float = 0.0, betas=(0.9, 0.999), eps: float = 1e-8,
                  grad_scale: float = 1.0, zero_grad: bool = True):
        """torch.optim.Adam(lr, weight_decay) semantics (parser.py:35-40) over the flat buffers: one kernel."""
        from . import _lib, ops
        stream = ops._prep(self.flat_p)
        lib = _lib.load()
        self.step_count += 1
        # torch.optim.Adam skips a parameter whose .grad is None (no weight decay, no moment decay, its own step count):
        # wsi_adam_step_masked reads the per-parameter flags on the device
        rc = lib.wsi_adam_step_masked(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.exp_avg.data_ptr(),
                                      self.exp_avg_sq.data_ptr(), self.numel, len(self.params), self.offs_dev.data_ptr(),
                                      self.flags_dev.data_ptr(), self.steps_dev.data_ptr(), self.corr_ws.data_ptr(), float(lr),
                                      float(betas[0]), float(betas[1]), float(eps), float(weight_decay), float(grad_scale),
                                      1 if zero_grad else 0, stream)
        _lib.check(rc, "wsi_adam_step_masked")


def train_step(model, graphs, labels: torch.Tensor, global_batch: int, optimizer, reducer: FlatGradAllReduce,
               loss_fn=torch.nn.functional.cross_entropy) -> torch.Tensor:
    """One data-parallel step on this rank's slides.  The local loss is the SUM over the local slides divided by the
    GLOBAL batch size, so the summed gradients equal those of the reference's mean CrossEntropyLoss over the whole
    batch (parser.py:182-183; trainer/train_gnn.py:68-71).  `graphs`: a packed HeteroGraph (hetero_graph.pack) or a
    list of graphs (legacy tuple branch, train_gnn.py:59-62).  -> the rank's loss contribution (detached)."""
    optimizer.zero_grad(set_to_none=True)
    if isinstance(graphs, (list, tuple)):
        logits = torch.cat([model(g) for g in graphs], 0)
    else:
        logits = model(graphs)
    loss = loss_fn(logits, labels, reduction="sum") / float(global_batch)
    loss.backward()
    reducer()
    optimizer.step()
    return loss.detach()


def flat_train_step(model, flat: FlatModel, micro_batches: Sequence, labels: Sequence[torch.Tensor], global_batch: int,
                    lr: float, weight_decay: float, loss_fn=torch.nn.functional.cross_entropy, events=None) -> torch.Tensor:
    """The same step on the flat buffers: `micro_batches` = this rank's slides as a list of packed HeteroGraphs (gradient
    accumulation bounds the activation memory), bucketed all-reduce overlapped with the last backward, fused Adam.
    events (optional): dict that receives CUDA events 'fwd' / 'bwd' / 'comm' / 'opt' lists of (start, end) pairs."""
    total = None

    def mark(kind, a, b):
        if events is not None:
            events.setdefault(kind, []).append((a, b))

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e
    timed = events is not None and labels[0].is_cuda
    flat.begin_step()
    for i, (G, y) in enumerate(zip(micro_batches, labels)):
        if i + 1 == len(micro_batches):
            flat.arm()
        t0 = ev() if timed else None
        logits = model(G)
        loss = loss_fn(logits, y, reduction="sum") / float(global_batch)
        t1 = ev() if timed else None
        loss.backward()
        if i + 1 < len(micro_batches):
            flat.fold()
        t2 = ev() if timed else None
        if timed:
            mark("fwd", t0, t1)
            mark("bwd", t1, t2)
        total = loss.detach() if total is None else total + loss.detach()
    t3 = ev() if timed else None
    flat.finish()
    t4 = ev() if timed else None
    flat.adam_step(lr, weight_decay)
    t5 = ev() if timed else None
    if timed:
        mark("comm", t3, t4)
        mark("opt", t4, t5)
    return total
