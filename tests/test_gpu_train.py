"""GPU parity of the TRAINING path: gradients of every parameter through the CUDA forward + backward kernels against
torch autograd on the CPU oracle (same parameters, same graph, same loss)."""
import pytest
import torch

import golden_util
import helpers
from wsi_hgnn_b200 import synthetic
from wsi_hgnn_b200.hetero_graph import pack

pytestmark = pytest.mark.gpu


def _grads(model, G, labels, device):
    model.zero_grad(set_to_none=True)
    if device == "cpu" and G.independent:
        from wsi_hgnn_b200.hetero_graph import unbatch
        logits = torch.cat([model(g) for g in unbatch(G)], 0)
    else:
        logits = model(G.to(device) if device != "cpu" else G)
    loss = torch.nn.functional.cross_entropy(logits, labels.to(logits.device))
    loss.backward()
    return loss.detach().cpu(), {k: p.grad.detach().cpu() for k, p in model.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("model,kw,n", [
    ("HEATNet4", dict(in_dim=64, hidden_dim=128, out_dim=3, n_layers=2, n_heads=4, dropuout=0.0), 700),
    ("HEATNet2", dict(in_dim=32, hidden_dim=256, out_dim=2, n_layers=1, n_heads=8, dropuout=0.0), 300),
    ("HEATNet4", dict(in_dim=64, hidden_dim=512, out_dim=2, n_layers=2, n_heads=4, dropuout=0.0), 1500),
    ("HEATNet2", dict(in_dim=32, hidden_dim=128, out_dim=2, n_layers=2, n_heads=4, dropuout=0.0, graph_pooling_type="max"), 400),
    ("HGT", dict(in_dim=48, hidden_dim=128, out_dim=2, n_layers=3, n_heads=4, use_norm=True), 600),
    ("HGT", dict(in_dim=32, hidden_dim=256, out_dim=3, n_layers=2, n_heads=8, use_norm=False, graph_pooling_type="sum"), 300),
])
def test_gradients_match_oracle_autograd(model, kw, n):
    gs = [synthetic.synth_slide_graph(n + 50 * i, kw["in_dim"], 3, 5, seed=20 + i, noise_edges=0.3) for i in range(2)]
    gs.append(synthetic.random_hetero_graph([60, 0, 40], 700, kw["in_dim"], seed=5, hub=120))     # empty type + hub
    G = pack(gs)
    labels = torch.tensor([0, 1, 1]) % kw["out_dim"]
    ours = helpers.build_ours(model, 3, kw)
    orc = helpers.build_oracle(model, 3, kw)
    golden_util.fill_params(ours, 99)
    orc.load_state_dict(ours.state_dict(), strict=True)
    ours = ours.cuda().train()
    orc = orc.double().train()
    for m_ in (ours, orc):                                  # HGTLayer hard-codes nn.Dropout(0.2) (models/HGT.py:34): off for parity
        for mod in m_.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
    for nt in G.ntypes:
        pass
    G64 = pack(gs)
    for nt in G64.ntypes:
        G64.nodes[nt].data["feat"] = G64.nodes[nt].data["feat"].double()
    for mod in orc.modules():
        if hasattr(mod, "e_linear"):
            mod.e_linear.float()                        # the reference casts sim to fp32 (models/HEATNet4.py:103)
    l_ref, g_ref = _grads(orc, G64, labels, "cpu")
    l_out, g_out = _grads(ours, G, labels, "cuda")
    assert abs(float(l_out) - float(l_ref)) < 1e-4 * max(1.0, abs(float(l_ref)))
    assert set(g_out) == set(g_ref), "parameters with gradients differ"
    worst = 0.0
    for k in g_ref:
        e = helpers.rel_err(g_out[k], g_ref[k])
        scale = float(g_ref[k].double().norm())
        if scale > 1e-9:
            worst = max(worst, e)
            assert e < 2e-3, f"grad of {k}: rel err {e:.3e}"
    assert worst > 0.0


def test_train_step_reduces_loss():
    """A few Adam steps through parallel.train_step (single rank) lower the loss; eval forward afterwards agrees with
    the training-mode forward (the cached weight packs are rebuilt after the in-place optimizer updates)."""
    from wsi_hgnn_b200.parallel import FlatGradAllReduce, train_step
    gs = [synthetic.synth_slide_graph(600, 32, 3, 5, seed=30 + i, noise_edges=0.2) for i in range(4)]
    G = pack(gs).to("cuda")
    labels = torch.tensor([0, 1, 0, 1], device="cuda")
    kw = dict(in_dim=32, hidden_dim=128, out_dim=2, n_layers=2, n_heads=4, dropuout=0.0)
    m = helpers.build_ours("HEATNet4", 3, kw)
    golden_util.fill_params(m, 3)
    m = m.cuda().train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    red = FlatGradAllReduce(m.parameters())
    losses = [float(train_step(m, G, labels, 4, opt, red)) for _ in range(8)]
    assert losses[-1] < losses[0]
    m.eval()
    with torch.no_grad():
        a = m(G)
    with torch.enable_grad():
        m.train()
        b = m(G).detach()
    assert helpers.rel_err(a, b) < 1e-4


def test_flat_model_step_equals_torch_adam():
    """parallel.FlatModel (flat parameter / gradient / moment buffers, gradients adopted from autograd and folded with
    multi-tensor adds, fused Adam kernel) == torch.optim.Adam on an identical model: three steps of two micro-batches
    each, parameters and the gradients exposed on p.grad compared."""
    import copy
    from wsi_hgnn_b200.parallel import FlatModel, flat_train_step
    gs = [synthetic.synth_slide_graph(500 + 50 * i, 32, 3, 5, seed=50 + i, noise_edges=0.2) for i in range(4)]
    packs = [pack(gs[:2]).to("cuda"), pack(gs[2:]).to("cuda")]
    labels = [torch.tensor([0, 1], device="cuda"), torch.tensor([1, 0], device="cuda")]
    kw = dict(in_dim=32, hidden_dim=128, out_dim=2, n_layers=2, n_heads=4, dropuout=0.0)
    ref = helpers.build_ours("HEATNet4", 3, kw)
    golden_util.fill_params(ref, 5)
    ref = ref.cuda().train()
    ours = copy.deepcopy(ref)
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3, weight_decay=5e-3)
    flat = FlatModel(ours)
    for step in range(3):
        opt.zero_grad(set_to_none=True)
        for G, y in zip(packs, labels):
            (torch.nn.functional.cross_entropy(ref(G), y, reduction="sum") / 4.0).backward()
        grads_ref = [p.grad.clone() if p.grad is not None else torch.zeros_like(p) for p in ref.parameters()]
        opt.step()
        # the flat step zeroes its gradient buffer inside the Adam kernel: look at the gradients through a hook-free path
        flat.begin_step()
        for i, (G, y) in enumerate(zip(packs, labels)):
            if i == 1:
                flat.arm()
            (torch.nn.functional.cross_entropy(ours(G), y, reduction="sum") / 4.0).backward()
            if i == 0:
                flat.fold()
        flat.finish()
        for (n, p), gr in zip(ours.named_parameters(), grads_ref):
            assert p.grad is not None and p.grad.data_ptr() >= flat.flat_g.data_ptr()
            assert helpers.rel_err(p.grad, gr) < 1e-4 or float(gr.abs().max()) < 1e-7, (step, n)
        flat.adam_step(1e-3, 5e-3)
    for (n, a), b in zip(ours.named_parameters(), ref.parameters()):
        assert helpers.rel_err(a, b) < 1e-4, n
    # and through the packaged step function
    l0 = float(flat_train_step(ours, flat, packs, labels, 4, 1e-3, 5e-3))
    l1 = float(flat_train_step(ours, flat, packs, labels, 4, 1e-3, 5e-3))
    assert l1 < l0 + 1e-3 and float(flat.flat_g.abs().max()) == 0.0       # gradients cleared by the fused step


def test_attn_bwd_two_pass_equals_atomic_mode():
    """wsi_hetero_attn_bwd: the two-pass mode (coefficients + source-major second kernel, no atomics) and the one-pass
    vector-atomic mode produce the same dK / dV / dQ / d e_linear (up to summation order: the atomics', and that of the
    one-sweep accumulation the two-pass mode uses for segments of more than two edges)."""
    from wsi_hgnn_b200 import ops
    D, H = 256, 4
    G = synthetic.synth_slide_graph(900, 32, 3, 6, seed=4, noise_edges=0.3).to("cuda")
    plan = G.plan()
    g = torch.Generator().manual_seed(0)
    kvq = torch.randn(plan.N, 3 * D, generator=g).cuda()
    d_agg = torch.randn(plan.N, D, generator=g).cuda()
    ew, eb = torch.tensor([0.7]).cuda(), torch.tensor([-0.1]).cuda()
    outs = []
    for transposed in (None, ops.transposed_edges(plan.rowptr, plan.e_src, plan.N)):
        dk, dv = torch.zeros(plan.N, D, device="cuda"), torch.zeros(plan.N, D, device="cuda")
        dq = torch.empty(plan.N, D, device="cuda")
        d_e = ops.hetero_attn_bwd(kvq[:, :D], kvq[:, D:2 * D], kvq[:, 2 * D:], plan.rowptr, plan.e_src, plan.e_sim, plan.e_rel,
                                  plan.node_inv_r, ew, eb, D, H, d_agg, dk, dv, dq, row_order=plan.rows_by_degree(),
                                  transposed=transposed)
        outs.append((dk, dv, dq, d_e))
    for a, b in zip(*outs):
        assert helpers.rel_err(a, b) < 1e-5
