"""Generate tests/golden/*.pt by executing the REFERENCE's own model files
(/root/reference/models/{HEATNet4,HEATNet2,HGT}.py and pooling/*.py, unmodified) on top of the
minimal DGL stand-in tests/dgl_shim.py.  Run in the build container (the reference is not present
on the GPU box); the outputs are committed fixtures.

    python tools/make_golden.py

Each fixture holds: the graph (HeteroGraph.state()), the model kind + ctor kwargs, the reference
parameters (refilled deterministically by tests/golden_util.fill_params; the fixture stores seed + checksum + shapes)
and the reference's logits in eval mode (fp32 and, for a tighter oracle check, fp64).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import dgl_shim  # noqa: E402
import golden_util  # noqa: E402
from wsi_hgnn_b200 import synthetic  # noqa: E402
from wsi_hgnn_b200.hetero_graph import batch, pack  # noqa: E402

# HEATNet4.py:240 hard-codes `.cuda()` for empty node types; on this CPU-only box make it a no-op.
torch.Tensor.cuda = lambda self, *a, **k: self


def edge_dict_for(n_types, etypes):
    # parser.py:125-134 ordering: for r for s for t
    cets = [(str(s), r, str(t)) for r in etypes for s in range(n_types) for t in range(n_types)]
    return {et: i for i, et in enumerate(cets)}


PARAM_SEED = 612

CASES = [
    # name, model, graph builder, ctor kwargs
    dict(name="heat4_rand_T3", model="HEATNet4",
         graph=lambda: synthetic.random_hetero_graph([20, 17, 11], 160, 24, seed=3),
         kw=dict(in_dim=24, hidden_dim=32, out_dim=2, n_layers=2, n_heads=4, dropuout=0.2, graph_pooling_type="mean")),
    dict(name="heat4_hub_T2_sum", model="HEATNet4",
         graph=lambda: synthetic.random_hetero_graph([30, 25], 120, 16, seed=4, hub=70),
         kw=dict(in_dim=16, hidden_dim=64, out_dim=3, n_layers=1, n_heads=4, dropuout=0.0, graph_pooling_type="sum")),
    dict(name="heat4_knn_T3_max", model="HEATNet4",
         graph=lambda: synthetic.synth_slide_graph(96, 32, 3, 5, seed=5, noise_edges=0.3),
         kw=dict(in_dim=32, hidden_dim=32, out_dim=2, n_layers=3, n_heads=2, dropuout=0.2, graph_pooling_type="max")),
    dict(name="heat4_emptytype_T4", model="HEATNet4",
         graph=lambda: synthetic.random_hetero_graph([25, 0, 18, 9], 140, 16, seed=6),
         kw=dict(in_dim=16, hidden_dim=32, out_dim=2, n_layers=2, n_heads=4, dropuout=0.2, graph_pooling_type="mean")),
    dict(name="heat2_rand_T3", model="HEATNet2",
         graph=lambda: synthetic.random_hetero_graph([22, 14, 19], 150, 24, seed=7),
         kw=dict(in_dim=24, hidden_dim=32, out_dim=4, n_layers=2, n_heads=4, dropuout=0.2, graph_pooling_type="mean")),
    dict(name="heat2_batch3_T2", model="HEATNet2",
         graph=lambda: _batched([synthetic.random_hetero_graph([12, 9], 90, 16, seed=s) for s in (8, 9, 10)]),
         kw=dict(in_dim=16, hidden_dim=32, out_dim=2, n_layers=2, n_heads=4, dropuout=0.2, graph_pooling_type="mean")),
    dict(name="hgt_rand_T3_norm", model="HGT",
         graph=lambda: synthetic.random_hetero_graph([18, 15, 13], 150, 24, seed=11),
         kw=dict(in_dim=24, hidden_dim=32, out_dim=2, n_layers=2, n_heads=4, use_norm=True, graph_pooling_type="mean")),
    dict(name="hgt_dk50_T2_nonorm_sum", model="HGT",
         graph=lambda: synthetic.random_hetero_graph([16, 12], 110, 20, seed=12),
         kw=dict(in_dim=20, hidden_dim=200, out_dim=2, n_layers=2, n_heads=4, use_norm=False, graph_pooling_type="sum")),
    # shapes that take the lane-grouped vector attention kernel (D % 128 == 0) and the tcgen05 typed GEMM
    dict(name="heat4_vec_D128_T3", model="HEATNet4",
         graph=lambda: synthetic.random_hetero_graph([70, 50, 40], 700, 64, seed=21, hub=45),
         kw=dict(in_dim=64, hidden_dim=128, out_dim=2, n_layers=2, n_heads=4, dropuout=0.2, graph_pooling_type="mean")),
    dict(name="heat4_vec_D512_T3_knn", model="HEATNet4",
         graph=lambda: synthetic.synth_slide_graph(300, 128, 3, 5, seed=22, noise_edges=0.25),
         kw=dict(in_dim=128, hidden_dim=512, out_dim=2, n_layers=3, n_heads=4, dropuout=0.2, graph_pooling_type="mean")),
    dict(name="heat2_vec_D256_H8_T2", model="HEATNet2",
         graph=lambda: synthetic.random_hetero_graph([90, 60], 600, 64, seed=23),
         kw=dict(in_dim=64, hidden_dim=256, out_dim=3, n_layers=2, n_heads=8, dropuout=0.2, graph_pooling_type="sum")),
    dict(name="hgt_D128_T3_norm", model="HGT",
         graph=lambda: synthetic.random_hetero_graph([60, 45, 30], 600, 64, seed=24, hub=40),
         kw=dict(in_dim=64, hidden_dim=128, out_dim=2, n_layers=3, n_heads=4, use_norm=True, graph_pooling_type="mean")),
    # the trainer's tuple branch (trainer/train_gnn.py:59-62): cat of independent forwards of graphs with
    # DIFFERENT relation sets / empty types, which pack() reproduces in one launch
    dict(name="heat4_pack3_T3", model="HEATNet4", independent=True,
         graph=lambda: [synthetic.random_hetero_graph([14, 9, 11], 60, 16, seed=31),
                        synthetic.random_hetero_graph([10, 0, 12], 25, 16, seed=32),
                        synthetic.random_hetero_graph([6, 7, 5], 12, 16, seed=33)],
         kw=dict(in_dim=16, hidden_dim=32, out_dim=2, n_layers=2, n_heads=4, dropuout=0.2, graph_pooling_type="mean")),
    # BASELINE.json configs[0], literally: 2 node types, 1k nodes, k=5 -> 5k edges, F = D = 64, H = 4, 1 layer
    dict(name="heat4_config1_T2", model="HEATNet4",
         graph=lambda: synthetic.synth_slide_graph(1000, 64, 2, 5, seed=0),
         kw=dict(in_dim=64, hidden_dim=64, out_dim=2, n_layers=1, n_heads=4, dropuout=0.2, graph_pooling_type="mean")),
    dict(name="heat2_pack2_T3_max", model="HEATNet2", independent=True,
         graph=lambda: [synthetic.random_hetero_graph([15, 0, 9], 70, 16, seed=41),
                        synthetic.random_hetero_graph([7, 12, 5], 40, 16, seed=42)],
         kw=dict(in_dim=16, hidden_dim=32, out_dim=3, n_layers=2, n_heads=4, dropuout=0.2, graph_pooling_type="max")),
    dict(name="hgt_pack2_T2", model="HGT", independent=True,
         graph=lambda: [synthetic.random_hetero_graph([12, 9], 50, 16, seed=34),
                        synthetic.random_hetero_graph([8, 11], 9, 16, seed=35)],
         kw=dict(in_dim=16, hidden_dim=32, out_dim=3, n_layers=2, n_heads=4, use_norm=True, graph_pooling_type="mean")),
    # round 2: the ESCA / COAD type count (T = 6: up to 72 relations, most of them a handful of edges)
    dict(name="hgt_T6_norm", model="HGT",
         graph=lambda: synthetic.random_hetero_graph([20, 15, 10, 8, 6, 4], 420, 24, seed=61, hub=30),
         kw=dict(in_dim=24, hidden_dim=64, out_dim=2, n_layers=3, n_heads=4, use_norm=True, graph_pooling_type="mean")),
    # round 2: big enough (>= 512 nodes and (dst, relation) segments, D % 128 == 0) for the tensor-core HGT schedule
    dict(name="hgt_tc_D128_T2_knn", model="HGT",
         graph=lambda: synthetic.synth_slide_graph(640, 24, 2, 6, seed=71, noise_edges=0.3),
         kw=dict(in_dim=24, hidden_dim=128, out_dim=2, n_layers=3, n_heads=4, use_norm=True, graph_pooling_type="mean")),
    # a relation that EXISTS with zero edges (what DropEdge / dgl.batch of sparse slides produce): it still counts in the
    # denominator of the cross-relation mean (multi_update_all(..., 'mean')), here in a dgl.batch-style batch of two
    dict(name="heat4_batch2_zero_edge_rel", model="HEATNet4",
         graph=lambda: _batched([_drop_relation(synthetic.random_hetero_graph([12, 10], 60, 16, seed=51), 1),
                                 _drop_relation(synthetic.random_hetero_graph([9, 14], 50, 16, seed=52), 1)]),
         kw=dict(in_dim=16, hidden_dim=32, out_dim=2, n_layers=2, n_heads=4, dropuout=0.2, graph_pooling_type="mean")),
]


def _drop_relation(g, index):
    """The same graph with every edge of its `index`-th canonical etype removed; the (now empty) relation stays."""
    from wsi_hgnn_b200.hetero_graph import HeteroGraph
    ce = g.canonical_etypes[index]
    edges = {c: (g._edges[c][0].clone(), g._edges[c][1].clone()) for c in g.canonical_etypes}
    edata = {c: {k: v.clone() for k, v in g._edata[c].items()} for c in g.canonical_etypes}
    edges[ce] = (edges[ce][0][:0], edges[ce][1][:0])
    edata[ce] = {k: v[:0] for k, v in edata[ce].items()}
    return HeteroGraph({nt: g.num_nodes(nt) for nt in g.ntypes}, edges, {nt: dict(g._ndata[nt]) for nt in g.ntypes}, edata)


def _batched(gs):
    # dgl.batch requires a common relation set; random graphs of this density have all of them.
    return batch(gs)


def main():
    mods = dgl_shim.load_reference_models("/root/reference")
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    only = set(sys.argv[1:])
    for case in CASES:
        if only and case["name"] not in only:
            continue
        G = case["graph"]()
        graphs = [G]
        if case.get("independent"):
            graphs = G
            G = pack(graphs)
        T = len(G.ntypes)
        node_dict = {str(i): i for i in range(T)}
        kw = dict(case["kw"])
        torch.manual_seed(611)
        if case["model"] == "HGT":
            ref = mods["HGT"].HGT(node_dict, edge_dict_for(T, ["neg", "pos"]), **kw)
        else:
            ref = getattr(mods[case["model"]], case["model"])(node_dict=node_dict, **kw)
        chk = golden_util.fill_params(ref, PARAM_SEED)
        ref.eval()
        outs = {}
        for dt in (torch.float32, torch.float64):
            m = ref.to(dt)
            for mod in m.modules():             # the reference casts sim to fp32 before e_linear (HEATNet4.py:103)
                if hasattr(mod, "e_linear"):
                    mod.e_linear.float()
            rows = []
            for g1 in graphs:
                sg = dgl_shim.shim_graph_from(g1)
                for nt in sg.ntypes:
                    sg._nframes[nt]["feat"] = sg._nframes[nt]["feat"].to(dt)
                with torch.no_grad():
                    rows.append(m(sg).detach().clone())
            outs[dt] = torch.cat(rows, 0)
        ref.to(torch.float32)
        fx = dict(name=case["name"], model=case["model"], kwargs=kw, graph=G.state(),
                  independent=bool(case.get("independent")),
                  param_seed=PARAM_SEED, param_checksum=chk,
                  param_shapes={k: tuple(v.shape) for k, v in ref.state_dict().items()},
                  logits_fp32=outs[torch.float32], logits_fp64=outs[torch.float64],
                  made_by="tools/make_golden.py: reference models on tests/dgl_shim.py")
        path = os.path.join(out_dir, case["name"] + ".pt")
        torch.save(fx, path)
        print(f"{case['name']:28s} logits {tuple(outs[torch.float32].shape)}  {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
