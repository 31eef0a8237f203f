"""Slide data sets over the flat format: the role of the reference's `GraphDataset` /
`TCGACancerTypingDataset` (data.py:66-290) without pickles or DGL.

The reference's data sets read a text file of graph paths (one per line), unpickle a DGLHeteroGraph per item, derive the
label from the path (TCGA barcode looked up in a list / mapping file) and apply the training transform when
`type_ == "train"` (data.py:116-117).  `FlatSlideDataset` keeps that contract - a list file of `*.wsiflat` paths, a
label function of the path, a transform on the training split only - but an item is a memory-mapped `FlatSlide` (host,
un-parsed) or, with `as_graph=True`, the HeteroGraph on `device`.  `convert_dgl_pickles` is the one-off converter to run
where DGL exists.
"""
import os
import pickle
from typing import Callable, Dict, List, Optional, Sequence, Union

import torch

from .hetero_graph import HeteroGraph
from .slide_io import FlatSlide


def tcga_barcode(path: str, n: int = 12) -> str:
    """`s[pos:pos + n]` from the first "TCGA" in the path (data.py:103-107: n = 16 for the normal lists, :264-265: n = 12
    for the typing mappings)."""
    pos = str(path).find("TCGA")
    if pos < 0:
        raise ValueError(f"no TCGA barcode in {path!r}")
    return str(path)[pos:pos + n]


def normal_list_label(normal_list: Sequence[str]) -> Callable[[str], int]:
    """label 0 if the slide's 16-character barcode is in the normal list else 1 (data.py:99-112)."""
    normal = set(normal_list)
    return lambda path: 0 if tcga_barcode(path, 16) in normal else 1


def mapping_label(mapping: Dict[str, Union[int, str]], classes: Optional[Dict[str, int]] = None) -> Callable[[str], int]:
    """label from a barcode -> class mapping (data.py:264-275); `classes` maps class names to ids, unknown names raise."""
    def f(path: str) -> int:
        lb = mapping[tcga_barcode(path, 12)]
        if classes is None:
            return int(lb)
        if lb not in classes:
            raise ValueError("Undefined label")
        return classes[lb]
    return f


class FlatSlideDataset(torch.utils.data.Dataset):
    def __init__(self, graph_path: Union[str, Sequence[str]], label_fn: Callable[[str], int], type_: str = "test",
                 transform: Optional[Callable[[HeteroGraph], HeteroGraph]] = None, as_graph: bool = False,
                 device: Union[str, torch.device] = "cpu", mmap: bool = True):
        if isinstance(graph_path, (str, os.PathLike)):
            with open(graph_path) as g:                              # a list file, one path per line (data.py:81-82)
                self.graph_paths: List[str] = [a.strip() for a in g.readlines() if a.strip()]
        else:
            self.graph_paths = [str(p) for p in graph_path]
        self.label_fn, self.type_, self.transform = label_fn, type_, transform
        self.as_graph, self.device, self.mmap = as_graph or transform is not None, torch.device(device), mmap

    def __len__(self) -> int:
        return len(self.graph_paths)

    def __getitem__(self, index: int):
        path = self.graph_paths[index]
        slide = FlatSlide.load(path, mmap=self.mmap)
        label = self.label_fn(path)
        if not self.as_graph:
            return slide, label
        g = slide.to_graph(self.device)
        if self.type_ == "train" and self.transform is not None:     # augmentation on the training split only (:116-117)
            g = self.transform(g)
        return g, label


def collate_pack(batch):
    """DataLoader collate_fn: the list of per-slide forwards of the reference's trainer (train_gnn.py:59-62) as one packed
    graph + a label tensor."""
    from .hetero_graph import pack
    graphs, labels = zip(*batch)
    return pack(list(graphs)), torch.tensor(labels, dtype=torch.long)


def convert_dgl_pickles(paths: Sequence[str], out_dir: str) -> List[str]:
    """One-off conversion of the reference's pickled DGLHeteroGraphs (get_graph.py:279-289) to `*.wsiflat` files; needs
    DGL importable (for unpickling).  -> the written paths."""
    os.makedirs(out_dir, exist_ok=True)
    out = []
    for p in paths:
        with open(p, "rb") as f:
            g = pickle.load(f)
        dst = os.path.join(out_dir, os.path.splitext(os.path.basename(p))[0] + ".wsiflat")
        FlatSlide.from_graph(HeteroGraph.from_dgl(g)).save(dst)
        out.append(dst)
    return out
