"""Data-parallel training over the GPUs of one box: slides are sharded over ranks (sharding.py), every rank runs
forward + backward on its own slides, and the ONE collective of the step is an all-reduce (sum) of a single flat fp32
gradient buffer over NCCL / NVLink.  The reference has no distributed code at all
(trainer/trainer.py:32-34 is single-device); this is the config-5 wrapper around its train_one_step
(trainer/train_gnn.py:55-79)."""
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class FlatGradAllReduce:
    """Packs the gradients of `params` into one flat buffer, all-reduces it once, unpacks in place."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        self.flat: Optional[torch.Tensor] = None

    def __call__(self):
        if not self.params:
            return
        dev, dt = self.params[0].device, torch.float32
        if self.flat is None or self.flat.device != dev:
            self.flat = torch.zeros(self.numel, dtype=dt, device=dev)
        off = 0
        for p in self.params:                           # parameters unused in the forward (HEATLayer.weight, attn.*)
            n = p.numel()                               # have no grad: they contribute zeros
            if p.grad is None:
                self.flat[off:off + n].zero_()
            else:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is not None:
                p.grad.copy_(self.flat[off:off + n].view_as(p.grad))
            off += n


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
