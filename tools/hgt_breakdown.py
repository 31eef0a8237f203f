"""Per-stage device times of one HGT layer of the tensor-core schedule at the config-3 shape (development)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import golden_util, helpers
    from wsi_hgnn_b200 import ops, synthetic
    from wsi_hgnn_b200.hetero_graph import pack
    from wsi_hgnn_b200.models import hgt as H
    prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    ops.set_matmul_precision(prec)
    sweep = [a for a in sys.argv[2:] if "=" in a and not a.startswith("fix:")]        # attention-only knob sweep: key=v1,v2,...
    for a in sys.argv[2:]:
        if a.startswith("fix:"):                          # fix:key=value - held for the whole run
            k_, v_ = a[4:].split("=")
            ops.dev_set(k_, int(v_))
    dev = torch.device("cuda", 0)
    T, k, F = 6, 6, 1024
    g = torch.Generator().manual_seed(99)
    sizes = torch.randint(4000, 12001, (16,), generator=g).tolist()
    graphs = [synthetic.device_slide_graph(n, F, T, k, seed=100 + i, device=dev, skew=True) for i, n in enumerate(sizes)]
    G = pack(graphs)
    kw = dict(in_dim=F, hidden_dim=512, out_dim=2, n_layers=4, n_heads=4, use_norm=True, graph_pooling_type="mean")
    model = helpers.build_ours("HGT", T, kw)
    golden_util.fill_params(model, 611)
    model = model.to(dev).eval()
    plan = G.plan()
    layer = model.gcs[0]
    D, Hh = 512, 4
    x = torch.randn(plan.N, D, device=dev)
    order = H._graph_type_order(plan, layer.node_dict)
    segs = plan.segments()
    grp = H._relation_groups(plan, layer.edge_dict, id(layer.edge_dict))
    opf = ops.matmul_opf()
    pk = layer._packed_tc(order, opf, dev)
    tpc = plan.type_ptr_c()
    S = segs["S"]
    st16 = opf != ops.OPF_BF16X3
    state = {}

    def s_xs(): state["xs"] = ops.to_operand(x, opf)
    def s_kv():
        if opf == ops.OPF_BF16:
            _, state["kv"] = ops.typed_linear_op(state["xs"], pk["w_kv"], pk["b_kv"], plan.type_ptr, 2 * D, want_y=False, want_op=True, type_ptr_c=tpc, opf=opf)
        else:
            state["kv"], _ = ops.typed_linear_op(state["xs"], pk["w_kv"], pk["b_kv"], plan.type_ptr, 2 * D, type_ptr_c=tpc, opf=opf)
    def s_q(): state["q"], _ = ops.typed_linear_op(state["xs"], pk["w_q"], pk["b_q"], plan.type_ptr, D, type_ptr_c=tpc, opf=opf)
    def s_qg(): state["qg"] = ops.gather_to_operand(state["q"], grp["dst_of_order"], opf)
    def s_qseg():
        a, b = ops.typed_linear_op(state["qg"], pk["w_att"], None, grp["rel_ptr"], D, type_ptr_c=grp["rel_ptr_c"], opf=opf, want_y=not st16, want_op=st16)
        state["qseg"] = b if st16 else a
    def s_attn():
        kv = state["kv"]
        sg = grp["seg_graph"]
        state["aggseg"] = ops.hetero_attn_work(kv[:, :D], kv[:, D:], state["qseg"], sg["work"], sg["e_src"], sg["e_sim"], sg["e_rel"], sg["inv"], sg["ew"], sg["eb"], D, Hh, op_out=True, opf=opf)
    def s_msg():
        a, b = ops.typed_linear_op(state["aggseg"], pk["w_msg"], None, grp["rel_ptr"], D, type_ptr_c=grp["rel_ptr_c"], opf=opf, want_y=not st16, want_op=st16)
        state["msg"] = b if st16 else a
    def s_comb(): _, state["aggs"] = ops.segment_combine(state["msg"], segs["row_seg_ptr"], plan.node_inv_r, plan.N, D, seg_pos=grp["seg_pos"], want_out=False, op_out=True, opf=opf)
    def s_alin(): state["out"], _ = ops.typed_linear_op(state["aggs"], pk["wa"], pk["ba"], plan.type_ptr, D, skip=pk["skip"], res=x, row_gate=plan.node_inv_r, type_ptr_c=tpc, opf=opf)
    def s_ln(): H._gated_layernorm(state["out"], pk["gamma"], pk["beta"], plan, tpc)
    stages = [("to_operand x", s_xs), ("K|V gemm", s_kv), ("Q gemm", s_q), ("gather q -> operand", s_qg), ("relation_att gemm", s_qseg),
              ("segment attention", s_attn), ("relation_msg gemm", s_msg), ("segment combine", s_comb), ("a_linear + skip", s_alin), ("layernorm", s_ln)]
    for _, f in stages:
        f()
    torch.cuda.synchronize()
    if sweep:
        deg = (grp["seg_graph"]["rowptr"][1:] - grp["seg_graph"]["rowptr"][:-1]).float()
        print(json.dumps({"seg_deg_mean": float(deg.mean()), "seg_deg_max": float(deg.max()), "n_items": grp["seg_graph"]["work"]["n_items"],
                          "n_split": grp["seg_graph"]["work"]["n_split"], "n_part": grp["seg_graph"]["work"]["n_part"]}))
        for sw in sweep:
            key, vals = sw.split("=")
            for v in vals.split(","):
                ops.dev_set(key, int(v))
                for _ in range(2):
                    s_attn()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(5):
                    s_attn()
                b.record()
                torch.cuda.synchronize()
                print(json.dumps({"knob": key, "value": int(v), "attn_us": a.elapsed_time(b) / 5 * 1e3}), flush=True)
            ops.dev_set(key, 0)
        return
    rec = {"precision": prec, "N": plan.N, "E": plan.E, "S": int(S), "R_nonempty": sum(1 for r in range(grp["R"]) if grp["rel_ptr"][r + 1] > grp["rel_ptr"][r])}
    tot = 0.0
    for name, f in stages:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            f()
        a.record()
        for _ in range(5):
            f()
        b.record()
        torch.cuda.synchronize()
        rec[name] = a.elapsed_time(b) / 5 * 1e3
        tot += rec[name]
    rec["sum_us"] = tot
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
