"""The edge builder of the reference's construct_graph/graph_constructor.py on the GPU.

* ``Hnsw``                 - same fit()/query() surface as reference graph_constructor.py:43-81, but EXACT
                             (brute-force L2 on the GPU, (distance, index) order) instead of nmslib's approximate
                             HNSW index; ``query_all`` returns every row's neighbours in one launch.
* ``GraphConstructor``     - ``construct_graph()`` of reference graph_constructor.py:256-303: k-NN edges
                             (query node -> neighbour, rank 0 dropped), Pearson edge type / ``sim``,
                             homogeneous -> heterogeneous conversion.  The CNN feature / node-type inference
                             wrappers of the reference (Hovernet_infer, KimiaNet_infer, EfficientNet_infer) are image
                             models upstream of the graph and out of scope: features and node types are passed in.
"""
from collections import OrderedDict
from typing import Optional, Sequence, Union

import numpy as np
import torch

from .. import ops
from ..hetero_graph import HeteroGraph, to_heterogeneous


def _as_device_features(X, device) -> torch.Tensor:
    if isinstance(X, np.ndarray):
        X = torch.from_numpy(np.ascontiguousarray(X))
    return X.to(device=device, dtype=torch.float32).contiguous()


class Hnsw:
    """Drop-in for the reference's ``Hnsw`` (graph_constructor.py:43-81).  ``index_params`` /
    ``query_params`` are accepted and ignored: the search is exact."""

    def __init__(self, space="cosinesimil", index_params=None, query_params=None, print_progress=True,
                 device="cuda"):
        if space != "l2":
            # the only space the reference ever instantiates is 'l2' (graph_constructor.py:226)
            raise NotImplementedError(f"space={space!r}: only 'l2' is implemented")
        self.space = space
        self.index_params = index_params
        self.query_params = query_params
        self.print_progress = print_progress
        self.device = torch.device(device)

    def fit(self, X):
        self.index_ = _as_device_features(X, self.device)
        self.index_params_ = self.index_params or {"M": 16, "post": 0, "efConstruction": 400}
        self.query_params_ = self.query_params or {"ef": 90}
        self._all = None
        return self

    def query_all(self, topn: int) -> torch.Tensor:
        """int32 [N, topn] neighbours of every indexed row (rank 0 = the row itself unless duplicated)."""
        if self._all is None or self._all.shape[1] != topn:
            self._all = ops.knn_topk(self.index_, topn)
        return self._all

    def query(self, vector, topn):
        """Neighbours of one vector (reference signature).  Prefer :meth:`query_all`."""
        v = _as_device_features(np.asarray(vector)[None] if not torch.is_tensor(vector) else vector[None], self.device)
        both = torch.cat([self.index_, v], 0)
        n = both.shape[0]
        nbr = ops.knn_topk(both, min(topn + 1, n), n - 1, n)[0]
        nbr = nbr[nbr != n - 1][:topn]              # drop the appended query row itself
        return nbr.cpu().numpy()


def construct_graph_arrays(features: torch.Tensor, radius: int):
    """(edge_index int64 [2, N*(radius-1)], edge_type uint8 [E], sim fp32 [E]) on the device, as
    construct_graph() assembles them (graph_constructor.py:262-282)."""
    n = features.shape[0]
    nbr = ops.knn_topk(features, radius)                      # [N, radius]; column 0 dropped (:270)
    src = torch.arange(n, device=features.device, dtype=torch.int64).repeat_interleave(radius - 1)   # :267
    dst = nbr[:, 1:].reshape(-1).to(torch.int64)
    sim, et = ops.edge_pearson(features, src, dst)
    return torch.stack([src, dst]), et, sim


class GraphConstructor:
    """reference graph_constructor.py:217-303 with the CNN stages replaced by given arrays.

    ``config`` needs ``radius`` and ``n_node_type`` (the keys construct_graph() reads, :222,295)."""

    def __init__(self, config: Union[OrderedDict, dict], features, node_type: Sequence[int], device="cuda"):
        self.config = config
        self.radius = config["radius"]
        self.device = torch.device(device)
        self.knn_model = Hnsw(space="l2", device=device)
        self.features = features
        self.node_type = node_type

    def construct_graph(self):
        feats = _as_device_features(self.features, self.device)
        ei, et, sim = construct_graph_arrays(feats, self.radius)
        ntype = torch.as_tensor(np.asarray(self.node_type), dtype=torch.int64, device=self.device)
        T = self.config["n_node_type"]
        het = to_heterogeneous(ei[0], ei[1], ntype, et.to(torch.int64), [str(t) for t in range(T)], ["neg", "pos"],
                               ndata={"feat": feats}, edata={"sim": sim})
        homo = HeteroGraph({"_N": feats.shape[0]}, {("_N", "_E", "_N"): (ei[0], ei[1])}, {"_N": {"feat": feats}})
        return het, homo, self.node_type
