"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes exercise the
shard partition and the all-gather-v of instance rows; the GPU counterpart is
test_gpu_multi.py (2 ranks, nccl)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bevy_firework_b200.distributed import all_gather_rows, shard_by_load, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # each rank "simulates" its shard of 7 spawners: rows tagged with spawner index
        mine = shard_range(7, world, rank)
        rows = []
        for s in mine:
            n = 3 + 2 * s  # ragged: different live counts per spawner
            r = torch.zeros((n, 16), dtype=torch.float32)
            r[:, 0] = s
            r[:, 1] = torch.arange(n)
            rows.append(r)
        local = torch.cat(rows) if rows else torch.zeros((0, 16))
        gathered, counts = all_gather_rows(local)
        np.save(os.path.join(out_dir, f"g{rank}.npy"), gathered.numpy())
        np.save(os.path.join(out_dir, f"c{rank}.npy"), np.array(counts))
        # empty contribution from one rank
        g2, c2 = all_gather_rows(local if rank == 0 else torch.zeros((0, 16)))
        assert c2[1] == 0 and g2.shape[0] == c2[0]
    finally:
        dist.destroy_process_group()


def test_shard_partition():
    assert [list(shard_range(512, 8, r)) for r in (0, 7)] == [list(range(0, 64)), list(range(448, 512))]
    covered = sorted(i for r in range(3) for i in shard_range(7, 3, r))
    assert covered == list(range(7))
    assert list(shard_range(2, 4, 3)) == []
    bins = shard_by_load([5, 1, 1, 1, 4, 4], 2)
    assert sorted(sum(bins, [])) == list(range(6))
    loads = [sum([5, 1, 1, 1, 4, 4][i] for i in b) for b in bins]
    assert abs(loads[0] - loads[1]) <= 1


def test_all_gather_v_two_ranks(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g0, g1 = np.load(tmp_path / "g0.npy"), np.load(tmp_path / "g1.npy")
    c0 = np.load(tmp_path / "c0.npy")
    assert (g0 == g1).all()                      # every rank holds the whole scene
    want_counts = [sum(3 + 2 * s for s in shard_range(7, 2, r)) for r in range(2)]
    assert list(c0) == want_counts
    # rank order, spawner order, Vec order are preserved
    expect = np.concatenate([[(s, i) for i in range(3 + 2 * s)] for s in range(7)])
    assert (g0[:, :2] == expect).all()
