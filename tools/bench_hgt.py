"""BASELINE config 3: TCGA-ESCA-shape batch of 16 graphs (4k-12k nodes each, k=6, T=6), 4-layer HGT + typed (nt) pooling
readout on one B200 (fp32 storage here; the reference's D=200 / d_k=50 configuration with --hidden 200).

    python tools/bench_hgt.py [--hidden 512] [--graphs 16] [--check 2]

Graphs are built ON THE GPU by the product's edge builder.  --check K compares the logits of the first K graphs with
the CPU oracle (rel err, north_star tolerance 1e-3).  Prints one JSON line.  Development / DESIGN.md numbers."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graphs", type=int, default=16)
    ap.add_argument("--hidden", type=int, default=512)
    ap.add_argument("--feat", type=int, default=1024)
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--check", type=int, default=0)
    ap.add_argument("--precision", default="bf16", help="bf16 (config 3 as stated: bf16 storage) | fp16 | bf16x3 (fp32-grade)")
    args = ap.parse_args()
    import golden_util
    import helpers
    from wsi_hgnn_b200 import ops, synthetic
    from wsi_hgnn_b200.construct_graph import GraphConstructor
    from wsi_hgnn_b200.hetero_graph import pack
    ops.set_matmul_precision(args.precision)

    dev = torch.device("cuda", 0)
    T, k = 6, 6
    g = torch.Generator().manual_seed(99)
    sizes = torch.randint(4000, 12001, (args.graphs,), generator=g).tolist()
    graphs = []
    t0 = time.perf_counter()
    for i, n in enumerate(sizes):
        feats, ntype = synthetic.synth_features(n, args.feat, T, seed=100 + i, skew=True)
        het, _, _ = GraphConstructor({"radius": k + 1, "n_node_type": T}, feats, ntype.numpy(), device=dev).construct_graph()
        graphs.append(het)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    G = pack(graphs)          # independent per-graph forwards (trainer/train_gnn.py:59-62): relation sets differ per slide
    kw = dict(in_dim=args.feat, hidden_dim=args.hidden, out_dim=2, n_layers=args.layers, n_heads=4, use_norm=True,
              graph_pooling_type="mean")
    model = helpers.build_ours("HGT", T, kw)
    golden_util.fill_params(model, 611)
    model = model.to(dev).eval()
    with torch.no_grad():
        for _ in range(3):
            out = model(G)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            out = model(G)
        b.record()
        torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    err = None
    if args.check > 0:
        orc = helpers.build_oracle("HGT", T, kw)
        orc.load_state_dict(model.state_dict(), strict=True)
        orc.eval()
        sub = pack(graphs[:args.check])
        with torch.no_grad():
            ref = helpers.run_oracle(orc, sub.to("cpu"), independent=True)
            got = model(sub)
        err = helpers.rel_err(got, ref)
    E, N, D = G.num_edges(), G.num_nodes(), args.hidden
    s_kv = 2 if args.precision == "bf16" else 4
    edge_bytes = (E * (2 * D * s_kv + 8) + N * (2 * D * 4 + 4)) * args.layers          # SURVEY 8(d) edge-phase bytes per layer
    print(json.dumps({"bench": "config3: HGT forward on a batch of ESCA-shape graphs", "graphs": args.graphs,
                      "precision": args.precision, "edge_phase_bytes_model": edge_bytes,
                      "edge_phase_model_share_of_forward_at_hbm_peak": edge_bytes / 6545.9e9 / (ms * 1e-3),
                      "nodes": G.num_nodes(), "edges": G.num_edges(), "hidden": args.hidden, "layers": args.layers,
                      "fwd_ms": ms, "edges_per_s": G.num_edges() / (ms * 1e-3), "graph_build_s": t_build,
                      "rel_err_vs_oracle_first_graphs": err, "logits0": out[0].cpu().tolist()}), flush=True)


if __name__ == "__main__":
    main()
