"""Multi-GPU plumbing: one process per GPU, spawners sharded across ranks, no collective on the
simulation path. The only exchange is the optional render extract: an
all-gather-v of the per-GPU ParticleInstance buffers, either through NCCL
(``all_gather_instances``) or fused into the pack kernel as peer stores over NVLink / NVSwitch
(``PeerGather``: the library's ``fw_gather_*`` exports; torch.distributed only carries the 88-byte
buffer handles once).

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


class PeerGather:
    """All-gather-v of the instance rows fused into the pack kernel: every rank maps every rank's
    gather buffer (CUDA IPC) and ``fw_gather_instances`` stores this rank's rows straight into its
    region of all of them, with device-side ready / landed flags instead of a host barrier.

        pg = PeerGather(engine, cap_rows_per_rank)      # collective (handle exchange)
        pg.issue()                                      # collective, asynchronous
        rows, counts = pg.result()                      # torch view of the gathered rows
    """

    def __init__(self, engine, cap_rows_per_rank: int, group=None):
        self.engine = engine
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        mine = engine.gather_create(self.world, self.rank, int(cap_rows_per_rank))
        handles: List[bytes] = [b""] * self.world
        dist.all_gather_object(handles, mine, group=group)
        engine.gather_connect(handles)
        self.group = group

    def issue(self):
        self.engine.gather_instances()

    def result(self) -> Tuple[torch.Tensor, List[int]]:
        """rows of all ranks concatenated in rank order (a copy; the regions themselves stay in the
        gather buffer at ``rank * stride`` rows for consumers that draw per region)"""
        ptr, counts, stride = self.engine.gather_result(self.world)
        dev = torch.device("cuda", self.engine.device)
        parts = [_device_view(ptr + r * stride * 64, counts[r], dev) for r in range(self.world)]
        return torch.cat(parts, dim=0), counts

    def close(self):
        self.engine.gather_destroy()


def _device_view(ptr: int, n_rows: int, device) -> torch.Tensor:
    """[n_rows, 16] float32 tensor aliasing library-owned device memory"""
    if n_rows == 0:
        return torch.empty((0, 16), dtype=torch.float32, device=device)

    class _Mem:
        __cuda_array_interface__ = {"shape": (n_rows, 16), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}

    return torch.as_tensor(_Mem(), device=device)
