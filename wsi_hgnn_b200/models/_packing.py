"""Host-side packing of the reference-shaped parameters into the layouts the kernels consume.

The modules keep the reference's parameter names and shapes (state_dict compatibility,
SURVEY.md §8b); the kernels want per-type stacks in the GRAPH's node-type order, K|V|Q fused
along the output dimension and, for the vector attention kernel, the lane-grouped column order.
Packs are cached and rebuilt when a parameter is modified in place (optimizer step, load_state_dict).
"""
from typing import Callable, Dict, Sequence, Tuple

import torch


def param_list(module, tag: str, fn):
    """Parameter list `fn()` of `module`, enumerated once (walking nn.Module.parameters() on every forward costs
    more host time than the launches it feeds).  The set of Parameter objects of a built model never changes."""
    cache = module.__dict__.setdefault("_plists", {})
    if tag not in cache:
        cache[tag] = list(fn())
    return cache[tag]


class PackCache:
    def __init__(self):
        self._store: Dict = {}

    def get(self, tag, params: Sequence[torch.Tensor], build: Callable[[], Tuple]):
        key = tuple((p.data_ptr(), p._version) for p in params)
        hit = self._store.get(tag)
        if hit is not None and hit[0] == key:
            return hit[1]
        with torch.no_grad():
            val = build()
        self._store[tag] = (key, val)
        return val

    def clear(self):
        self._store.clear()


class _Permute(torch.autograd.Function):
    """x.index_select(dim, perm) for a PERMUTATION perm: the backward is the gather by the inverse permutation
    (torch's generic advanced-indexing backward is a sort + atomic scatter, ~100x slower for these weight stacks)."""

    @staticmethod
    def forward(ctx, x, perm, dim):
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(perm.numel(), device=perm.device)
        ctx.inv, ctx.dim = inv, dim
        return x.index_select(dim, perm)

    @staticmethod
    def backward(ctx, g):
        return g.index_select(ctx.dim, ctx.inv), None, None


def stack_linears(linears, order: Sequence[int], row_perm=None, col_perm=None):
    """[T, n_out, K] weight stack and [T, n_out] bias stack of `linears[i] for i in order`.
    The (optional) output-row / input-column permutation is applied ONCE to the stacked tensors: permuting every
    nn.Linear separately cost ~7 T small index kernels per layer and step on the training path."""
    w = torch.stack([linears[i].weight for i in order])
    b = torch.stack([linears[i].bias for i in order])
    if row_perm is not None:
        w, b = _Permute.apply(w, row_perm, 1), _Permute.apply(b, row_perm, 1)
    if col_perm is not None:
        w = _Permute.apply(w, col_perm, 2)
    return w.contiguous(), b.contiguous()
