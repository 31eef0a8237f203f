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


def test_hgt_relation_sorted_segment_structures():
    """Host structures of the HGT tensor-core schedule (models/hgt._relation_groups): the (dst, relation) segments in
    relation order - work items, inverse position, relation pointers - describe exactly the segments of the plan."""
    from wsi_hgnn_b200 import synthetic
    from wsi_hgnn_b200.models.hgt import _relation_groups
    G = synthetic.random_hetero_graph([40, 30, 25], 400, 8, seed=5)
    plan = G.plan()
    segs = plan.segments()
    S = int(segs["S"])
    names = [str(t) for t in range(3)]
    edge_dict = {(s, r, d): i for i, (r, s, d) in enumerate((r, s, d) for r in ("neg", "pos") for s in names for d in names)}
    grp = _relation_groups(plan, edge_dict, "t")
    order = grp["order"].long()
    sp = segs["seg_ptr"].long()
    items = grp["items"].long()
    assert items.shape == (S, 4) and torch.equal(items[:, 0], torch.arange(S)) and bool((items[:, 3] == -1).all())
    assert torch.equal(items[:, 1], sp[:-1][order]) and torch.equal(items[:, 2], sp[1:][order])
    # relation ids are non-decreasing along the sorted order and rel_ptr delimits them
    rel_sorted = grp["seg_rel_sorted"].long()
    assert bool((rel_sorted[1:] >= rel_sorted[:-1]).all())
    rp = grp["rel_ptr"]
    assert rp[0] == 0 and rp[-1] == S and len(rp) == len(edge_dict) + 1
    for r in range(len(edge_dict)):
        assert bool((rel_sorted[rp[r]:rp[r + 1]] == r).all())
    # seg_pos is the inverse of the order; the dst row of every sorted segment is the segment's dst
    pos = grp["seg_pos"].long()
    assert torch.equal(pos[order], torch.arange(S)) and torch.equal(grp["dst_of_order"].long(), segs["seg_dst"].long()[order])
    assert grp["seg_graph"] is None                      # the work list of the segment graph is a device structure
    # every edge belongs to exactly one segment and the segment lengths survive the reordering
    assert int((items[:, 2] - items[:, 1]).sum()) == plan.E
