"""Multi-GPU plumbing: one process per GPU, spawners sharded across ranks, no collective on the
simulation path. The only exchange is the optional render extract: an
all-gather-v of the per-GPU ParticleInstance buffers over NCCL (NVLink / NVSwitch).

``torch.distributed`` is used purely as plumbing; with the ``gloo`` backend the same code runs
on CPU tensors (tests/test_distributed_gloo.py).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_spawners: int, world_size: int, rank: int) -> range:
    """contiguous block partition: spawner i -> rank i // ceil(S / G)."""
    per = -(-n_spawners // world_size)
    lo = min(rank * per, n_spawners)
    return range(lo, min(lo + per, n_spawners))


def shard_by_load(loads: Sequence[float], world_size: int) -> List[List[int]]:
    """greedy longest-processing-time partition by expected live count (rate x lifetime)."""
    order = sorted(range(len(loads)), key=lambda i: -loads[i])
    bins: List[List[int]] = [[] for _ in range(world_size)]
    totals = [0.0] * world_size
    for i in order:
        r = min(range(world_size), key=lambda k: totals[k])
        bins[r].append(i)
        totals[r] += loads[i]
    return [sorted(b) for b in bins]


def all_gather_rows(local_rows: torch.Tensor, group=None) -> Tuple[torch.Tensor, List[int]]:
    """all-gather-v of ``[n_local, 16]`` float32 instance rows: every rank receives the rows of
    all ranks concatenated in rank order, plus the per-rank counts. Counts first (tiny
    all_gather), then one padded ``all_gather_into_tensor`` so NCCL sees equal-sized chunks."""
    world = dist.get_world_size(group)
    n_local = torch.tensor([local_rows.shape[0]], dtype=torch.int64, device=local_rows.device)
    counts_t = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts_t, n_local, group=group)
    counts = [int(c.item()) for c in counts_t]
    width = local_rows.shape[1]
    n_max = max(max(counts), 1)
    padded = local_rows
    if local_rows.shape[0] != n_max:
        padded = torch.zeros((n_max, width), dtype=local_rows.dtype, device=local_rows.device)
        padded[: local_rows.shape[0]] = local_rows
    gathered = torch.empty((world * n_max, width), dtype=local_rows.dtype, device=local_rows.device)
    if dist.get_backend(group) == "gloo":
        chunks = list(gathered.view(world, n_max, width).unbind(0))
        dist.all_gather(chunks, padded.contiguous(), group=group)
    else:
        dist.all_gather_into_tensor(gathered, padded.contiguous(), group=group)
    parts = [gathered[r * n_max: r * n_max + counts[r]] for r in range(world)]
    return torch.cat(parts, dim=0), counts


def all_gather_instances(engine, group=None, slack_rows: int = 1 << 16) -> Tuple[torch.Tensor, List[int]]:
    """the whole scene's ParticleInstance rows on every GPU: ``fw_pack_instances_device`` into a
    torch buffer on this rank's device, then the all-gather-v above."""
    cap = engine.total_live() + slack_rows
    buf = torch.empty((cap, 16), dtype=torch.float32, device=f"cuda:{engine.device}")
    n = engine.pack_instances_device(buf.data_ptr(), cap)
    return all_gather_rows(buf[:n], group)
