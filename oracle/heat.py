"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of the reference HEAT trunk and the HEATNet4 / HEATNet2 readouts, structured
like the reference (per-relation loop, K/Q/V re-projected per relation, stack->mean) so it
doubles as the "reference-structured" CPU baseline.

* OracleHEATLayer  follows reference models/HEATNet4.py:49-138 (== models/HEATNet2.py:24-113)
* OracleHEATNet4   follows reference models/HEATNet4.py:141-247
* OracleHEATNet2   follows reference models/HEATNet2.py:116-196

Parameter names/shapes equal the reference's so one state_dict drives reference, oracle and
the CUDA modules.  The graph argument is duck-typed: ntypes, canonical_etypes,
edges(etype=), nodes[nt].data['feat'], edata['sim'], batch_num_nodes(nt).
"""
import math

import torch
import torch.nn as nn

from . import primitives as P


def _pool_op(name):
    if name not in ("mean", "sum", "max"):
        # 'att' (GlobalAttentionPooling) cannot take ntype= and raises in the reference too
        raise NotImplementedError(name)
    return name


class OracleHEATLayer(nn.Module):
    def __init__(self, in_size, out_size, node_dict, n_heads, dropout=0.2):
        super().__init__()
        self.weight = nn.Linear(in_size, out_size)           # models/HEATNet4.py:54 - created, never used
        self.out_size = out_size
        self.node_dict = node_dict
        T = len(node_dict)
        self.n_heads = n_heads
        self.d_k = out_size // n_heads
        self.sqrt_dk = math.sqrt(self.d_k)
        mk = lambda i, o: nn.ModuleList([nn.Linear(i, o) for _ in range(T)])
        self.k_linears = mk(in_size, out_size)
        self.q_linears = mk(in_size, out_size)
        self.v_linears = mk(in_size, out_size)
        self.a_linears = mk(out_size, out_size)
        self.e_linear = nn.Linear(1, 1)                       # models/HEATNet4.py:74
        self.skip = nn.Parameter(torch.ones(T))               # models/HEATNet4.py:76
        self.drop = nn.Dropout(dropout)

    def forward(self, G, feat_dict, sim_dict):
        H, dk = self.n_heads, self.d_k
        per_dst = {}
        for ce in G.canonical_etypes:                         # models/HEATNet4.py:91
            s_t, _, d_t = ce
            src, dst = G.edges(etype=ce)
            n_dst = feat_dict[d_t].shape[0]
            k = self.k_linears[self.node_dict[s_t]](feat_dict[s_t]).view(-1, H, dk)   # :100
            v = self.v_linears[self.node_dict[s_t]](feat_dict[s_t]).view(-1, H, dk)   # :101
            q = self.q_linears[self.node_dict[d_t]](feat_dict[d_t]).view(-1, H, dk)   # :102
            ea = self.e_linear(sim_dict[ce].view(-1, 1).type(torch.float32))          # :103
            t = P.v_dot_u(q, k, src, dst)                                             # :109
            score = t.sum(-1) * ea / self.sqrt_dk                                     # :111
            a = P.edge_softmax(score, dst, n_dst)                                     # :113
            m = P.u_mul_e_sum(v, a.unsqueeze(-1), src, dst, n_dst)                    # :114,118-119
            per_dst.setdefault(d_t, []).append(m)
        new_h = {}
        for nt in G.ntypes:                                                           # :122
            n_id = self.node_dict[nt]
            alpha = torch.sigmoid(self.skip[n_id])
            if nt not in per_dst:                                                     # KeyError path :129-133
                new_h[nt] = feat_dict[nt]
                continue
            t = P.cross_reduce_mean(per_dst[nt]).view(-1, self.out_size)
            trans = self.drop(self.a_linears[n_id](t))                                # :134
            new_h[nt] = trans * alpha + feat_dict[nt] * (1 - alpha)                   # :135
        return new_h


def _input_projection(model, G, h, act=None):
    out = {}
    for nt in G.ntypes:
        x = G.nodes[nt].data["feat"] if h is None else h[nt]
        y = model.adapt_ws[model.node_dict[nt]](x)
        out[nt] = act(y) if act is not None else y
    return out


def _sim_dict(G):
    sim = G.edata["sim"]
    if not isinstance(sim, dict):
        sim = {G.canonical_etypes[0]: sim}
    return sim


class OracleHEATNet4(nn.Module):
    def __init__(self, in_dim, hidden_dim, out_dim, n_layers, n_heads, node_dict, dropuout, graph_pooling_type="mean"):
        super().__init__()
        self.node_dict = node_dict
        self.n_layers = n_layers
        self.pool = _pool_op(graph_pooling_type)
        self.linears_prediction = nn.ModuleDict({k: nn.Linear(hidden_dim, 256) for k in node_dict})
        self.adapt_ws = nn.ModuleList([nn.Linear(in_dim, hidden_dim) for _ in node_dict])
        self.gcs = nn.ModuleList([OracleHEATLayer(hidden_dim, hidden_dim, node_dict, n_heads, dropuout)
                                  for _ in range(n_layers)])
        # attn.{k}.op: Conv1d(256,1,1,bias=False); LinearAttentionBlock == identity on its first
        # argument (softmax over an axis of length 1, models/HEATNet4.py:26-37).
        self.attn = nn.ModuleDict({k: _AttnParams() for k in node_dict})
        self.head_2 = nn.Linear(256 * len(node_dict), 256)
        self.head_1 = nn.Linear(256, 64)
        self.head = nn.Linear(64, out_dim)

    def forward(self, G, h=None, return_embeddings=False):
        h = _input_projection(self, G, h)                                             # :198-206
        sim = _sim_dict(G)                                                            # :209-210
        for i in range(self.n_layers):
            h = self.gcs[i](G, h, sim)                                                # :213-214
        parts = []
        for nt in G.ntypes:   # h.items() iterates in G.ntypes order                   # :216-240
            if h[nt].shape[0] > 0:
                pooled = P.segment_readout(h[nt], G.batch_num_nodes(nt), self.pool)
                parts.append(self.linears_prediction[nt](pooled))
            else:
                parts.append(torch.zeros(1, 256))
        g = torch.cat(parts, dim=1)
        g = self.head(self.head_1(self.head_2(g)))                                    # :242-245
        return (g, h) if return_embeddings else g


class _AttnParams(nn.Module):
    def __init__(self):
        super().__init__()
        self.op = nn.Conv1d(256, 1, kernel_size=1, padding=0, bias=False)


class OracleHEATNet2(nn.Module):
    def __init__(self, in_dim, hidden_dim, out_dim, n_layers, n_heads, node_dict, dropuout, graph_pooling_type="mean"):
        super().__init__()
        self.node_dict = node_dict
        self.n_layers = n_layers
        self.pool = _pool_op(graph_pooling_type)
        self.linears_prediction = nn.ModuleDict({k: nn.Linear(hidden_dim, out_dim) for k in node_dict})
        self.adapt_ws = nn.ModuleList([nn.Linear(in_dim, hidden_dim) for _ in node_dict])
        self.gcs = nn.ModuleList([OracleHEATLayer(hidden_dim, hidden_dim, node_dict, n_heads, dropuout)
                                  for _ in range(n_layers)])

    def forward(self, G, h=None, return_embeddings=False):
        h = _input_projection(self, G, h)                                             # HEATNet2.py:162-170
        sim = _sim_dict(G)
        for i in range(self.n_layers):
            h = self.gcs[i](G, h, sim)                                                # :178-179
        hg = 0
        for nt in G.ntypes:                                                           # :181-194
            if h[nt].shape[0] > 0:
                pooled = P.segment_readout(h[nt], G.batch_num_nodes(nt), self.pool)
                hg = hg + self.linears_prediction[nt](pooled)
        return (hg, h) if return_embeddings else hg
