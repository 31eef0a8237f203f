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


def stack_linears(linears, order: Sequence[int], row_perm=None, col_perm=None):
    """[T, n_out, K] weight stack and [T, n_out] bias stack of `linears[i] for i in order`."""
    ws, bs = [], []
    for i in order:
        w, b = linears[i].weight, linears[i].bias
        if row_perm is not None:
            w, b = w[row_perm], b[row_perm]
        if col_perm is not None:
            w = w[:, col_perm]
        ws.append(w)
        bs.append(b)
    return torch.stack(ws).contiguous(), torch.stack(bs).contiguous()
