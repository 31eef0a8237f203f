"""Typed graph readout with the reference's pooling/{avg,sum,max}_pooling.py call contract:
``forward(graph, feat_dict, ntype) -> [B, D]`` (dgl.readout.{mean,sum,max}_nodes(..., ntype=))."""
import torch
import torch.nn as nn

from .. import ops


class _TypedPooling(nn.Module):
    op = "sum"

    def forward(self, graph, feat, ntype=None):
        # reference pooling/avg_pooling.py:11-19: graph.ndata['h'] = feat; readout = mean_nodes(graph, 'h', ntype=ntype)
        plan = graph.plan()
        if isinstance(feat, dict):
            if ntype is None:
                if len(feat) != 1:
                    raise ValueError("ntype is required when several node types carry features")
                ntype = next(iter(feat))
            x = feat[ntype]
        else:
            if ntype is None:
                if len(plan.ntypes) != 1:
                    raise ValueError("ntype is required for a heterogeneous graph")
                ntype = plan.ntypes[0]
            x = feat
        t = plan.ntypes.index(ntype)
        B = plan.B
        seg = plan.seg_ptr[t * B:(t + 1) * B + 1] - plan.type_ptr[t]
        return ops.segment_pool(x.contiguous(), seg.contiguous(), B, self.op)


class AvgPooling(_TypedPooling):
    """reference pooling/avg_pooling.py:6-19"""
    op = "mean"


class SumPooling(_TypedPooling):
    """reference pooling/sum_pooling.py:6-18"""
    op = "sum"


class MaxPooling(_TypedPooling):
    """reference pooling/max_pooling.py:6-19"""
    op = "max"


class NTPooling(nn.Module):
    """reference pooling/nt_pooling.py:4-10 is a stub (`pass`); kept as such."""

    def forward(self, graph, feat, ntype=None):
        return None


__all__ = ["AvgPooling", "SumPooling", "MaxPooling", "NTPooling"]
