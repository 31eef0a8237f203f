"""Flat-format data set: list file, label from the TCGA barcode, transform on the training split only."""
import os

import pytest
import torch

from wsi_hgnn_b200 import synthetic
from wsi_hgnn_b200.data import FlatSlideDataset, collate_pack, mapping_label, normal_list_label, tcga_barcode
from wsi_hgnn_b200.slide_io import FlatSlide
from wsi_hgnn_b200.transforms import DropEdge


def _write(tmp_path):
    names = ["TCGA-AA-0001-01A-x", "TCGA-BB-0002-11A-y", "TCGA-CC-0003-01A-z"]
    paths, graphs = [], []
    for i, n in enumerate(names):
        g = synthetic.random_hetero_graph([20 + i, 10], 60 + 10 * i, 6, seed=i)
        p = os.path.join(tmp_path, n + ".wsiflat")
        FlatSlide.from_graph(g).save(p)
        paths.append(p)
        graphs.append(g)
    lst = os.path.join(tmp_path, "graphs.txt")
    open(lst, "w").write("\n".join(paths) + "\n")
    return lst, paths, graphs


def test_labels_from_barcodes():
    assert tcga_barcode("/d/TCGA-AA-0001-01A-x.wsiflat", 16) == "TCGA-AA-0001-01A" and tcga_barcode("TCGA-AA-0001-zz") == "TCGA-AA-0001"
    with pytest.raises(ValueError):
        tcga_barcode("/d/slide_17.wsiflat")
    f = normal_list_label(["TCGA-BB-0002-11A"])
    assert f("/x/TCGA-BB-0002-11A-y.wsiflat") == 0 and f("/x/TCGA-AA-0001-01A-x.wsiflat") == 1
    m = mapping_label({"TCGA-AA-0001": "Infiltrating Ductal Carcinoma", "TCGA-CC-0003": "other"},
                      {"Infiltrating Ductal Carcinoma": 0, "Infiltrating Lobular Carcinoma": 1})
    assert m("TCGA-AA-0001-01A") == 0
    with pytest.raises(ValueError, match="Undefined label"):
        m("TCGA-CC-0003-01A")
    assert mapping_label({"TCGA-AA-0001": "3"})("TCGA-AA-0001-01A") == 3


def test_dataset_items_and_train_only_transform(tmp_path):
    lst, paths, graphs = _write(tmp_path)
    label = normal_list_label(["TCGA-BB-0002-11A"])
    ds = FlatSlideDataset(lst, label)
    assert len(ds) == 3
    slide, y = ds[1]
    assert isinstance(slide, FlatSlide) and y == 0 and slide.num_edges() == graphs[1].num_edges()
    drop_all = DropEdge(1.0)
    test = FlatSlideDataset(paths, label, type_="test", transform=drop_all)
    train = FlatSlideDataset(paths, label, type_="train", transform=drop_all)
    g_test, _ = test[0]
    g_train, y0 = train[0]
    assert g_test.num_edges() == graphs[0].num_edges() and g_train.num_edges() == 0 and y0 == 1
    assert torch.equal(g_test.nodes["0"].data["feat"], graphs[0].nodes["0"].data["feat"])
    G, ys = collate_pack([test[i] for i in range(3)])
    assert G.batch_size == 3 and ys.tolist() == [1, 0, 1] and G.num_nodes() == sum(g.num_nodes() for g in graphs)
