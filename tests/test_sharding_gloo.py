"""Multi-rank host logic on CPU (gloo, world_size 2): LPT slide sharding + logits gather == single-process order.
The per-slide forward stand-in here is the CPU oracle (tests may use it); on GPUs each rank runs the CUDA path."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wsi_hgnn_b200 import synthetic
from wsi_hgnn_b200.sharding import gather_logits, lpt_assign, shard_slides


def test_lpt_assign_balances_and_covers():
    costs = [20000, 2100, 19000, 2500, 8000, 7000, 3000, 16000, 2000, 12000]
    for world in (1, 2, 3, 4, 8):
        owned = lpt_assign(costs, world)
        assert sorted(i for o in owned for i in o) == list(range(len(costs)))
        loads = [sum(costs[i] for i in o) for o in owned]
        assert max(loads) <= sum(costs) / world + max(costs)          # LPT bound
    two = lpt_assign(costs, 2)
    assert abs(sum(costs[i] for i in two[0]) - sum(costs[i] for i in two[1])) <= 0.1 * sum(costs)
    assert lpt_assign([], 4) == [[], [], [], []]


def _graphs():
    return [synthetic.random_hetero_graph([30 + 7 * i, 20 + 3 * i], 100 + 60 * i, 16, seed=40 + i) for i in range(7)]


def _model():
    import helpers
    import golden_util
    m = helpers.build_oracle("HEATNet4", 2, dict(in_dim=16, hidden_dim=32, out_dim=3, n_layers=2, n_heads=4, dropuout=0.0))
    golden_util.fill_params(m, 7)
    return m.eval()


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        graphs = _graphs()
        mine, gs = shard_slides(graphs, rank, world)
        m = _model()
        with torch.no_grad():
            local = torch.cat([m(g) for g in gs], 0) if gs else torch.zeros(0, 3)
        full = gather_logits(local, mine, len(graphs))
        q.put((rank, mine, full))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_inference_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m = _model()
    with torch.no_grad():
        ref = torch.cat([m(g) for g in _graphs()], 0)
    owned = sorted(i for _, mine, _ in res for i in mine)
    assert owned == list(range(7))
    for _, _, full in res:
        assert torch.allclose(full, ref, rtol=0, atol=0)


def _dp_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from wsi_hgnn_b200.parallel import FlatGradAllReduce
        torch.manual_seed(0)
        m = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
        frozen = torch.nn.Linear(2, 2)                                   # a parameter that never gets a grad
        params = list(m.parameters()) + list(frozen.parameters())
        g = torch.Generator().manual_seed(1)
        X, y = torch.randn(8, 6, generator=g), torch.randint(0, 3, (8,), generator=g)
        mine = list(range(rank, 8, world))                               # this rank's samples
        loss = torch.nn.functional.cross_entropy(m(X[mine]), y[mine], reduction="sum") / 8.0
        loss.backward()
        FlatGradAllReduce(params)()
        q.put((rank, [p.grad.clone() for p in m.parameters()]))
    finally:
        dist.destroy_process_group()


def test_flat_grad_allreduce_equals_full_batch_gradient():
    """sum over ranks of d(sum local CE / B_global) == d(mean CE over the whole batch) (reference loss, parser.py:182-183)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    g = torch.Generator().manual_seed(1)
    X, y = torch.randn(8, 6, generator=g), torch.randint(0, 3, (8,), generator=g)
    torch.nn.functional.cross_entropy(m(X), y).backward()
    for _, grads in res:
        for a, p in zip(grads, m.parameters()):
            assert torch.allclose(a, p.grad, rtol=1e-5, atol=1e-7)


def _flat_model_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from wsi_hgnn_b200.parallel import FlatModel
        torch.manual_seed(0)

        class Net(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.a = torch.nn.ModuleList([torch.nn.Linear(6, 5) for _ in range(3)])      # stacked per "type" below
                self.b = torch.nn.Linear(5, 3)
                self.unused = torch.nn.Linear(2, 2)                                          # never receives a gradient

            def forward(self, x):
                w = torch.stack([l.weight for l in self.a])                                  # grads arrive as views of d(stack)
                bb = torch.stack([l.bias for l in self.a])
                h = torch.tanh(torch.einsum("nk,tok->no", x, w) + bb.sum(0))
                return self.b(h)

        m = Net()
        flat = FlatModel(m, bucket_mb=1e-4)                               # tiny buckets: several of them, launched from the hooks
        g = torch.Generator().manual_seed(1)
        X, y = torch.randn(8, 6, generator=g), torch.randint(0, 3, (8,), generator=g)
        mine = list(range(rank, 8, world))
        halves = [mine[:len(mine) // 2], mine[len(mine) // 2:]]           # two micro-batches
        out = []
        for step in range(2):                                             # step 2 launches buckets early (expected sets known)
            flat.begin_step()
            for i, idx in enumerate(halves):
                if i + 1 == len(halves):
                    flat.arm()
                (torch.nn.functional.cross_entropy(m(X[idx]), y[idx], reduction="sum") / 8.0).backward()
                if i + 1 < len(halves):
                    flat.fold()
            flat.finish()
            out.append(([p.grad.tolist() for p in m.parameters()], [bool(f) for f in flat.flags_dev.tolist()], len(flat.buckets)))   # plain lists: no shared-memory handles in the queue
            flat.zero_grad()
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_flat_model_buckets_adopted_gradients_two_ranks():
    """parallel.FlatModel over 2 gloo ranks: gradients adopted from autograd (p.grad None during the step), folded per
    bucket with multi-tensor adds, bucketed all-reduce from the hooks on the second step; p.grad after finish() == the
    full-batch gradient; the parameter without a gradient is flagged inactive on every rank (torch.optim.Adam skips it)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_flat_model_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    a = torch.nn.ModuleList([torch.nn.Linear(6, 5) for _ in range(3)])
    b = torch.nn.Linear(5, 3)
    g = torch.Generator().manual_seed(1)
    X, y = torch.randn(8, 6, generator=g), torch.randint(0, 3, (8,), generator=g)
    w, bb = torch.stack([l.weight for l in a]), torch.stack([l.bias for l in a])
    torch.nn.functional.cross_entropy(b(torch.tanh(torch.einsum("nk,tok->no", X, w) + bb.sum(0))), y).backward()
    ref = [p.grad for p in list(a.parameters()) + list(b.parameters())]
    for _, steps in res:
        for grads, active, n_buckets in steps:
            assert n_buckets > 1
            grads = [torch.tensor(x) for x in grads]
            for got, want in zip(grads[:len(ref)], ref):
                assert torch.allclose(got, want, rtol=1e-5, atol=1e-7)
            assert all(float(x.abs().max()) == 0.0 for x in grads[len(ref):])       # the unused layer
            assert active[:len(ref)] == [True] * len(ref) and active[len(ref):] == [False, False]


def test_flat_model_single_process_micro_batches():
    """FlatModel without a process group (CPU tensors): two micro-batches folded into the flat buffer == the summed
    gradient; p.grad are views of the flat buffer afterwards; a parameter without a gradient stays inactive and zero."""
    from wsi_hgnn_b200.parallel import FlatModel
    torch.manual_seed(3)
    m = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Tanh(), torch.nn.Linear(4, 2))
    unused = torch.nn.Linear(3, 3)
    holder = torch.nn.ModuleList([m, unused])
    ref = [p.detach().clone() for p in m.parameters()]
    flat = FlatModel(holder, bucket_mb=1e-4)
    for p, r in zip(m.parameters(), ref):
        assert torch.equal(p, r) and p.data_ptr() >= flat.flat_p.data_ptr()        # parameters moved into the flat buffer
    X = torch.randn(6, 5)
    flat.begin_step()
    for i, idx in enumerate(([0, 1, 2], [3, 4, 5])):
        if i == 1:
            flat.arm()
        m(X[idx]).pow(2).sum().backward()
        if i == 0:
            assert all(p.grad is not None and p.grad.data_ptr() != v.data_ptr() for p, v in zip(m.parameters(), flat.views))
            flat.fold()
            assert all(p.grad is None for p in m.parameters())
    flat.finish()
    m2 = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Tanh(), torch.nn.Linear(4, 2))
    m2.load_state_dict({k: v.clone() for k, v in m.state_dict().items()})
    m2(X).pow(2).sum().backward()
    for (p, v), q in zip(zip(m.parameters(), flat.views[:4]), m2.parameters()):
        assert p.grad is v and torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-6)
    assert flat._active == [True] * 4 + [False, False]
    assert all(float(p.grad.abs().max()) == 0.0 for p in unused.parameters())
