"""Host-side weight packing (pure torch, CPU): the stacked single permutation equals permuting every nn.Linear on its
own, forward and backward (the gradients must land on the reference-shaped parameters)."""
import torch
import torch.nn as nn

from wsi_hgnn_b200.models._packing import PackCache, stack_linears


def test_stack_linears_permutation_and_grads():
    torch.manual_seed(0)
    T, n_out, K = 3, 8, 6
    lins = nn.ModuleList([nn.Linear(K, n_out) for _ in range(T)])
    order = [2, 0, 1]
    rp = torch.randperm(n_out)
    cp = torch.randperm(K)
    w, b = stack_linears(lins, order, row_perm=rp, col_perm=cp)
    for j, i in enumerate(order):
        assert torch.equal(w[j], lins[i].weight[rp][:, cp])
        assert torch.equal(b[j], lins[i].bias[rp])
    gw, gb = torch.randn_like(w), torch.randn_like(b)
    (w * gw).sum().backward(retain_graph=True)
    (b * gb).sum().backward()
    inv_r, inv_c = torch.argsort(rp), torch.argsort(cp)
    for j, i in enumerate(order):
        assert torch.allclose(lins[i].weight.grad, gw[j][inv_r][:, inv_c])
        assert torch.allclose(lins[i].bias.grad, gb[j][inv_r])
    w2, b2 = stack_linears(lins, order)
    assert torch.equal(w2[0], lins[2].weight) and torch.equal(b2[1], lins[0].bias)


def test_pack_cache_rebuilds_on_in_place_update():
    lin = nn.Linear(4, 4)
    cache, calls = PackCache(), []

    def build():
        calls.append(1)
        return (lin.weight.detach().clone(),)

    params = list(lin.parameters())
    a = cache.get("k", params, build)
    b = cache.get("k", params, build)
    assert a is b and len(calls) == 1
    with torch.no_grad():
        lin.weight.add_(1.0)                                   # optimizer step / load_state_dict bump the version
    c = cache.get("k", params, build)
    assert len(calls) == 2 and torch.equal(c[0], lin.weight)
