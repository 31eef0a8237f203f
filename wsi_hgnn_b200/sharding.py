"""Data-parallel sharding of slides over the GPUs of one box.

Slides are independent units - the reference itself treats a hetero batch as independent per-graph forwards
(trainer/train_gnn.py:59-62) - so the forward needs NO data-path collective: every rank (one process per GPU) runs
the slides it owns and the logits are gathered once at the end.  Slide sizes vary by 10x (2k-20k nodes), so the
assignment is greedy longest-processing-time on the edge count, not round-robin on the slide count.
"""
from typing import List, Sequence

import torch


def lpt_assign(costs: Sequence[float], world: int) -> List[List[int]]:
    """Greedy LPT: slides in decreasing cost order, each to the currently lightest rank.
    -> per-rank lists of slide indices (each list in increasing index order).  Deterministic."""
    if world < 1:
        raise ValueError("world must be >= 1")
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * world
    owned: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owned[r].append(i)
        load[r] += float(costs[i])
    return [sorted(o) for o in owned]


def shard_slides(graphs: Sequence, rank: int, world: int):
    """(indices, graphs) of the slides rank `rank` owns; cost = number of edges (+ nodes, for edge-free slides)."""
    costs = [g.num_edges() + 0.1 * g.num_nodes() for g in graphs]
    mine = lpt_assign(costs, world)[rank]
    return mine, [graphs[i] for i in mine]


def gather_logits(local: torch.Tensor, mine: Sequence[int], n_total: int, group=None) -> torch.Tensor:
    """All ranks' per-slide logits [len(mine), C] -> [n_total, C] in the original slide order, on every rank.
    One all_gather of a padded block of logits and one of the int64 slide indices (the only collectives of the inference path)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    C = int(local.shape[1]) if local.dim() == 2 else 0
    cnt = torch.tensor([len(mine)], dtype=torch.int64, device=local.device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    mx = int(max(int(c) for c in cnts))
    pad = torch.zeros((mx, C), dtype=local.dtype, device=local.device)
    pad[:len(mine)] = local
    idx = torch.full((mx,), -1, dtype=torch.int64, device=local.device)      # slide indices travel as int64 (exact for any
    idx[:len(mine)] = torch.tensor(list(mine), dtype=torch.int64, device=local.device)   # logits dtype, fp16 / bf16 included)
    blocks = [torch.zeros_like(pad) for _ in range(world)]
    idxs = [torch.zeros_like(idx) for _ in range(world)]
    dist.all_gather(blocks, pad, group=group)
    dist.all_gather(idxs, idx, group=group)
    out = torch.zeros((n_total, C), dtype=local.dtype, device=local.device)
    for b, ix, c in zip(blocks, idxs, cnts):
        k = int(c)
        if k:
            out[ix[:k]] = b[:k]
    return out


def bind_to_gpu_numa(device_index: int) -> dict:
    """Pin the calling process to the CPU cores of the NUMA node its GPU hangs off (one process per GPU): pinned host
    buffers allocated afterwards are first-touched on that node, so the H2D copies of 8 ranks do not all pull from one
    socket's memory.  Best effort - returns what was done ({} when /sys does not expose the topology)."""
    import os
    import torch
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return {}
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus), "pci": bdf}
    except Exception:                 # noqa: BLE001 - best effort: no CUDA device, no /sys topology, restricted affinity ...
        return {}
