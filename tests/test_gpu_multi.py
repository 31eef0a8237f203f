"""Multi-GPU paths over NCCL (needs >= 2 GPUs on the box; skipped otherwise): the node-sharded single-slide forward and
the sharded edge builder must reproduce the single-GPU results."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("kv_wire,tol", [("fp32", 2e-5), ("auto", 6e-4)])
@pytest.mark.parametrize("world", [2, 4])
def test_node_sharded_forward_and_edge_builder_over_nccl(world, kv_wire, tol):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tools", "bench_node_sharded.py"), "--nodes", "20000", "--feat", "256", "--hidden", "256",
           "--steps", "3", "--check", "--kv-wire", kv_wire]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["n_gpus"] == world and sum(d["rows_per_rank"]) == 20000
    assert d["logits_identical_on_all_ranks"] is True
    assert d["edge_index_bit_exact_vs_single"] is True              # bit-exact edge_index, sim and edge type
    assert d["rel_err_vs_unsharded"] < tol          # fp32 wire: the unsharded arithmetic; 16-bit wire: + one fp16 rounding of K|V
