"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Pure-torch CPU restatement of the DGL primitives the reference's hot path calls
([DGL-mem]: DGL is not installed here; semantics restated from DGL's documented behaviour).

* v_dot_u            - reference models/HEATNet4.py:109, models/HGT.py:99
* edge_softmax       - reference models/HEATNet4.py:113, models/HGT.py:101 (norm_by='dst')
* u_mul_e -> sum     - reference models/HEATNet4.py:118-119, models/HGT.py:105-106
* cross_reducer mean - same lines; stack(...).mean(0) over ALL relations into a dst type
* {mean,sum,max}_nodes - reference pooling/avg_pooling.py:15-17, sum_pooling.py:14-16, max_pooling.py:15-17
"""
from typing import List

import torch


def v_dot_u(q: torch.Tensor, k: torch.Tensor, src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """fn.v_dot_u('q','k','t'): t[e,h,0] = <q[dst e,h,:], k[src e,h,:]>; keeps a trailing dim of 1."""
    return (q.index_select(0, dst) * k.index_select(0, src)).sum(-1, keepdim=True)


def edge_softmax(score: torch.Tensor, dst: torch.Tensor, num_dst: int) -> torch.Tensor:
    """dgl.nn.edge_softmax(sub_graph, score) with the default norm_by='dst': softmax over the
    in-edges of each dst node of THIS relation, per head, max-subtracted."""
    if score.shape[0] == 0:
        return score.clone()
    idx = dst.view(-1, *([1] * (score.dim() - 1))).expand_as(score)
    mx = torch.full((num_dst,) + tuple(score.shape[1:]), float("-inf"), dtype=score.dtype)
    mx = mx.scatter_reduce(0, idx, score, reduce="amax", include_self=True)
    ex = torch.exp(score - mx.index_select(0, dst))
    den = torch.zeros((num_dst,) + tuple(score.shape[1:]), dtype=score.dtype).index_add_(0, dst, ex)
    return ex / den.index_select(0, dst)


def u_mul_e_sum(v: torch.Tensor, a: torch.Tensor, src: torch.Tensor, dst: torch.Tensor, num_dst: int) -> torch.Tensor:
    """update_all(fn.u_mul_e('v','t','m'), fn.sum('m','t')): t[d] = sum_{e->d} v[src e] * a[e] (zero-initialised)."""
    out = torch.zeros((num_dst,) + tuple(v.shape[1:]), dtype=v.dtype)
    if src.numel():
        out.index_add_(0, dst, v.index_select(0, src) * a)
    return out


def cross_reduce_mean(frames: List[torch.Tensor]) -> torch.Tensor:
    """multi_update_all(..., cross_reducer='mean'): a single relation is returned as is,
    otherwise torch.stack(frames).mean(0) - relations that delivered no message to a node
    still count in the denominator."""
    if len(frames) == 1:
        return frames[0]
    return torch.stack(frames, 0).mean(0)


def segment_readout(x: torch.Tensor, seglen: torch.Tensor, op: str) -> torch.Tensor:
    """dgl.readout.{mean,sum,max}_nodes(graph, 'h', ntype=...) -> [B, D]; empty segment -> 0."""
    B = int(seglen.numel())
    out = torch.zeros((B,) + tuple(x.shape[1:]), dtype=x.dtype)
    off = 0
    for b in range(B):
        n = int(seglen[b])
        if n > 0:
            seg = x[off:off + n]
            if op == "sum":
                out[b] = seg.sum(0)
            elif op == "mean":
                out[b] = seg.sum(0) / max(n, 1)
            elif op == "max":
                out[b] = seg.max(0).values
            else:
                raise NotImplementedError(op)
        off += n
    return out
