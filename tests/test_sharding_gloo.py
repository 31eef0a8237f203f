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
