"""A minimal stand-in for the DGL API the reference's hot path touches, so that the reference's
OWN model files (models/HEATNet4.py, HEATNet2.py, HGT.py, pooling/*.py) can be executed
unmodified in a container without DGL.  TEST INFRASTRUCTURE ONLY.

It is deliberately written as literal per-node / per-edge Python loops (a different
formulation from oracle/primitives.py, which is vectorised) so that agreement between
"reference-on-shim" and the oracle is a real cross-check of the message-passing semantics.
[DGL-mem]: semantics restated from DGL's documented behaviour; DGL itself is absent.

Install with ``install()`` BEFORE importing reference modules.
"""
import contextlib
import sys
import types

import torch


# ----------------------------------------------------------------------------- dgl.function
class _Msg:
    def __init__(self, kind, a, b, out):
        self.kind, self.a, self.b, self.out = kind, a, b, out


class _Red:
    def __init__(self, kind, msg, out):
        self.kind, self.msg, self.out = kind, msg, out


def _make_function_module():
    m = types.ModuleType("dgl.function")
    m.v_dot_u = lambda a, b, out: _Msg("v_dot_u", a, b, out)
    m.u_mul_e = lambda a, b, out: _Msg("u_mul_e", a, b, out)
    m.sum = lambda msg, out: _Red("sum", msg, out)
    return m


# ----------------------------------------------------------------------------- graph
class _TypeView:
    def __init__(self, frame):
        self.data = frame


class _Nodes:
    def __init__(self, g):
        self._g = g

    def __getitem__(self, nt):
        return _TypeView(self._g._nframes[nt])


class _MultiFrame:
    """G.edata / G.ndata on a graph with several types: get -> dict by type, set <- dict by type."""

    def __init__(self, frames, keys):
        self._frames, self._keys = frames, keys

    def __getitem__(self, name):
        if len(self._keys) == 1:
            return self._frames[self._keys[0]][name]
        return {k: self._frames[k][name] for k in self._keys if name in self._frames[k]}

    def __setitem__(self, name, val):
        if isinstance(val, dict):
            for k, v in val.items():
                self._frames[k][name] = v
        else:
            assert len(self._keys) == 1
            self._frames[self._keys[0]][name] = val


class _PopFrame(dict):
    pass


class _RelGraph:
    def __init__(self, g, ce):
        self._g, self.ce = g, ce
        self.srcdata = g._nframes[ce[0]]
        self.dstdata = g._nframes[ce[2]]
        self.edata = g._eframes[ce]

    def edges(self):
        return self._g._edges[self.ce]

    def num_dst(self):
        return self._g._num[self.ce[2]]

    def apply_edges(self, msg):
        assert msg.kind == "v_dot_u"
        src, dst = self.edges()
        q, k = self.dstdata[msg.a], self.srcdata[msg.b]
        rows = []
        for e in range(src.numel()):
            rows.append((q[int(dst[e])] * k[int(src[e])]).sum(-1, keepdim=True))
        self.edata[msg.out] = torch.stack(rows) if rows else q.new_zeros((0,) + tuple(q.shape[1:-1]) + (1,))


def edge_softmax(sub, score, norm_by="dst"):
    assert norm_by == "dst"
    src, dst = sub.edges()
    out = torch.empty_like(score)
    for v in range(sub.num_dst()):
        idx = torch.nonzero(dst == v).reshape(-1)
        if idx.numel():
            out[idx] = torch.softmax(score[idx], dim=0)
    return out


class ShimHeteroGraph:
    def __init__(self, num_nodes, edges, batch_num_nodes=None):
        self.ntypes = sorted(num_nodes)
        self.canonical_etypes = sorted(edges)
        self._num = dict(num_nodes)
        self._edges = {ce: (torch.as_tensor(s, dtype=torch.int64), torch.as_tensor(d, dtype=torch.int64))
                       for ce, (s, d) in edges.items()}
        self._nframes = {nt: _PopFrame() for nt in self.ntypes}
        self._eframes = {ce: _PopFrame() for ce in self.canonical_etypes}
        self._bnn = batch_num_nodes or {nt: [self._num[nt]] for nt in self.ntypes}

    @property
    def nodes(self):
        return _Nodes(self)

    @property
    def edata(self):
        return _MultiFrame(self._eframes, self.canonical_etypes)

    @property
    def ndata(self):
        return _MultiFrame(self._nframes, self.ntypes)

    def __getitem__(self, ce):
        return _RelGraph(self, tuple(ce))

    def edges(self, etype=None):
        return self._edges[tuple(etype)]

    def num_nodes(self, nt=None):
        return self._num[nt] if nt is not None else sum(self._num.values())

    def batch_num_nodes(self, nt):
        return torch.tensor(self._bnn[nt], dtype=torch.int64)

    @contextlib.contextmanager
    def local_scope(self):
        ns = {k: dict(v) for k, v in self._nframes.items()}
        es = {k: dict(v) for k, v in self._eframes.items()}
        try:
            yield
        finally:
            for k in self._nframes:
                self._nframes[k].clear()
                self._nframes[k].update(ns[k])
            for k in self._eframes:
                self._eframes[k].clear()
                self._eframes[k].update(es[k])

    def to(self, device):
        return self

    def multi_update_all(self, etype_dict, cross_reducer):
        assert cross_reducer == "mean"
        outs = {}
        out_name = None
        for ce, (msg, red) in etype_dict.items():
            assert msg.kind == "u_mul_e" and red.kind == "sum"
            ce = tuple(ce)
            src, dst = self._edges[ce]
            v = self._nframes[ce[0]][msg.a]
            a = self._eframes[ce][msg.b]
            acc = v.new_zeros((self._num[ce[2]],) + tuple(v.shape[1:]))
            for e in range(src.numel()):
                acc[int(dst[e])] = acc[int(dst[e])] + v[int(src[e])] * a[e]
            outs.setdefault(ce[2], []).append(acc)
            out_name = red.out
        for nt, frames in outs.items():
            self._nframes[nt][out_name] = frames[0] if len(frames) == 1 else torch.stack(frames).mean(0)


def _readout(op):
    def fn(graph, feat, ntype=None):
        x = graph._nframes[ntype][feat]
        rows, off = [], 0
        for n in graph._bnn[ntype]:
            seg = x[off:off + n]
            if n == 0:
                rows.append(x.new_zeros(x.shape[1:]))
            elif op == "mean":
                rows.append(seg.sum(0) / n)
            elif op == "sum":
                rows.append(seg.sum(0))
            else:
                rows.append(seg.max(0).values)
            off += n
        return torch.stack(rows)
    return fn


def install():
    """Register the shim as ``dgl`` (+ the sub-modules the reference imports) in sys.modules."""
    if "dgl" in sys.modules and not getattr(sys.modules["dgl"], "_is_wsi_shim", False):
        raise RuntimeError("a real dgl is already imported")
    dgl = types.ModuleType("dgl")
    dgl._is_wsi_shim = True
    dgl.DGLGraph = ShimHeteroGraph
    dgl.function = _make_function_module()
    nn_mod = types.ModuleType("dgl.nn")
    nn_mod.edge_softmax = edge_softmax
    pt = types.ModuleType("dgl.nn.pytorch")
    glob = types.ModuleType("dgl.nn.pytorch.glob")

    class GlobalAttentionPooling(torch.nn.Module):
        def __init__(self, gate_nn):
            super().__init__()
            self.gate_nn = gate_nn

    class MaxPooling(torch.nn.Module):       # DGL's homogeneous MaxPooling: no ntype=
        pass

    glob.GlobalAttentionPooling = GlobalAttentionPooling
    glob.MaxPooling = MaxPooling
    pt.glob = glob
    nn_mod.pytorch = pt
    dgl.nn = nn_mod
    ro = types.ModuleType("dgl.readout")
    ro.mean_nodes, ro.sum_nodes, ro.max_nodes = _readout("mean"), _readout("sum"), _readout("max")
    dgl.readout = ro
    sys.modules.update({"dgl": dgl, "dgl.function": dgl.function, "dgl.nn": nn_mod, "dgl.nn.pytorch": pt,
                        "dgl.nn.pytorch.glob": glob, "dgl.readout": ro})
    return dgl


def load_reference_models(ref_root="/root/reference"):
    """Import the reference's model files by path (models/__init__.py is broken: it imports a missing HAN)."""
    import importlib.util
    install()
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)          # so that `from pooling import ...` resolves to the reference's package
    mods = {}
    for name in ("HEATNet4", "HEATNet2", "HGT"):
        spec = importlib.util.spec_from_file_location(f"_wsi_ref_{name}", f"{ref_root}/models/{name}.py")
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods[name] = m
    return mods


def shim_graph_from(G):
    """Build a ShimHeteroGraph from our HeteroGraph (cpu)."""
    g = ShimHeteroGraph({nt: G.num_nodes(nt) for nt in G.ntypes},
                        {ce: G.edges(etype=ce) for ce in G.canonical_etypes},
                        {nt: [int(x) for x in G.batch_num_nodes(nt)] for nt in G.ntypes})
    for nt in G.ntypes:
        for k, v in G.nodes[nt].data.items():
            g._nframes[nt][k] = v
    for ce in G.canonical_etypes:
        for k, v in G[ce].edata.items():
            g._eframes[ce][k] = v
    return g
