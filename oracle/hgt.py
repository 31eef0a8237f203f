"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of the reference HGT: OracleHGTLayer follows reference models/HGT.py:21-127,
OracleHGT follows models/HGT.py:130-209.  Same parameter names/shapes as the reference.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import primitives as P
from .heat import _input_projection, _pool_op


class OracleHGTLayer(nn.Module):
    def __init__(self, in_dim, out_dim, node_dict, edge_dict, n_heads, dropout=0.2, use_norm=False):
        super().__init__()
        self.out_dim = out_dim
        self.node_dict = node_dict
        self.edge_dict = edge_dict
        T, R = len(node_dict), len(edge_dict)
        self.n_heads = n_heads
        self.d_k = out_dim // n_heads
        self.sqrt_dk = math.sqrt(self.d_k)
        self.use_norm = use_norm
        mk = lambda i, o: nn.ModuleList([nn.Linear(i, o) for _ in range(T)])
        self.k_linears = mk(in_dim, out_dim)
        self.q_linears = mk(in_dim, out_dim)
        self.v_linears = mk(in_dim, out_dim)
        self.a_linears = mk(out_dim, out_dim)
        self.norms = nn.ModuleList([nn.LayerNorm(out_dim) for _ in range(T)] if use_norm else [])
        self.relation_pri = nn.Parameter(torch.ones(R, n_heads))                       # HGT.py:59
        self.relation_att = nn.Parameter(torch.empty(R, n_heads, self.d_k, self.d_k))  # :60
        self.relation_msg = nn.Parameter(torch.empty(R, n_heads, self.d_k, self.d_k))  # :61
        self.skip = nn.Parameter(torch.ones(T))
        self.drop = nn.Dropout(dropout)
        nn.init.xavier_uniform_(self.relation_att)
        nn.init.xavier_uniform_(self.relation_msg)

    def forward(self, G, h):
        H, dk = self.n_heads, self.d_k
        per_dst = {}
        for ce in G.canonical_etypes:                                                  # :75
            if ce not in self.edge_dict:
                raise KeyError(ce)                                                     # :86 would raise too
            s_t, _, d_t = ce
            src, dst = G.edges(etype=ce)
            n_dst = h[d_t].shape[0]
            k = self.k_linears[self.node_dict[s_t]](h[s_t]).view(-1, H, dk)            # :82
            v = self.v_linears[self.node_dict[s_t]](h[s_t]).view(-1, H, dk)            # :83
            q = self.q_linears[self.node_dict[d_t]](h[d_t]).view(-1, H, dk)            # :84
            e_id = self.edge_dict[ce]
            k = torch.einsum("bij,ijk->bik", k, self.relation_att[e_id])               # :92
            v = torch.einsum("bij,ijk->bik", v, self.relation_msg[e_id])               # :93
            t = P.v_dot_u(q, k, src, dst)                                              # :99
            score = t.sum(-1) * self.relation_pri[e_id] / self.sqrt_dk                 # :100
            a = P.edge_softmax(score, dst, n_dst)                                      # :101
            per_dst.setdefault(d_t, []).append(P.u_mul_e_sum(v, a.unsqueeze(-1), src, dst, n_dst))  # :105-106
        new_h = {}
        for nt in G.ntypes:                                                            # :109
            n_id = self.node_dict[nt]
            alpha = torch.sigmoid(self.skip[n_id])
            if nt not in per_dst:                                                      # :118-120
                new_h[nt] = h[nt]
                continue
            t = P.cross_reduce_mean(per_dst[nt]).view(-1, self.out_dim)
            trans = self.drop(self.a_linears[n_id](t))                                 # :121
            trans = trans * alpha + h[nt] * (1 - alpha)                                # :122
            new_h[nt] = self.norms[n_id](trans) if self.use_norm else trans            # :123-126
        return new_h


class OracleHGT(nn.Module):
    def __init__(self, node_dict, edge_dict, in_dim, hidden_dim, out_dim, n_layers, n_heads,
                 use_norm=True, graph_pooling_type="mean"):
        super().__init__()
        self.node_dict = node_dict
        self.edge_dict = edge_dict
        self.n_layers = n_layers
        self.pool = _pool_op(graph_pooling_type)
        self.adapt_ws = nn.ModuleList([nn.Linear(in_dim, hidden_dim) for _ in node_dict])
        self.gcs = nn.ModuleList([OracleHGTLayer(hidden_dim, hidden_dim, node_dict, edge_dict, n_heads,
                                                 use_norm=use_norm) for _ in range(n_layers)])
        self.out = nn.Linear(hidden_dim, out_dim)                                      # :150 - unused
        self.linears_prediction = nn.ModuleDict({
            k: nn.ModuleList([nn.Linear(hidden_dim, out_dim) for _ in range(n_layers + 1)]) for k in node_dict})

    def forward(self, G, h=None, return_embeddings=False):
        h = _input_projection(self, G, h, act=F.gelu)                                  # :176-184
        hg = 0
        for i in range(self.n_layers):                                                 # :189-199
            for nt in G.ntypes:
                if h[nt].shape[0] > 0:
                    pooled = P.segment_readout(h[nt], G.batch_num_nodes(nt), self.pool)
                    hg = hg + self.linears_prediction[nt][i](pooled)
            h = self.gcs[i](G, h)
        return (hg, h) if return_embeddings else hg
