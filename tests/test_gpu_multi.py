"""Two ranks, one GPU each (nccl): spawners sharded by rank, no collective while simulating,
all-gather-v of the instance rows for a render extract. Needs >= 2 GPUs (gpurun --gpus 2)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
DT = float(np.float32(1.0) / np.float32(60.0))
N_SPAWNERS = 8
FRAMES = 70


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _simulate(engine, indices):
    from bevy_firework_b200._native import frame_input
    from bevy_firework_b200.workloads import grid_positions, stress_spawner

    sp = stress_spawner(rate=2500.0)
    pos = grid_positions(N_SPAWNERS)
    inputs = []
    for i in indices:
        ps, nt, es, ne = sp.pods()
        engine.spawner_reset(100 + i, ps, nt, es, ne, True)
        inputs.append(frame_input(100 + i, pos[i]))
    for _ in range(FRAMES):
        engine.frame(DT, inputs)


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    from bevy_firework_b200._native import Engine
    from bevy_firework_b200.distributed import all_gather_instances, shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        eng = Engine(device=rank, seed=0x00F12E00)
        _simulate(eng, shard_range(N_SPAWNERS, world, rank))
        rows, counts = all_gather_instances(eng)
        np.save(os.path.join(out_dir, f"rows{rank}.npy"), rows.cpu().numpy())
        np.save(os.path.join(out_dir, f"counts{rank}.npy"), np.array(counts))
        # the same exchange fused into the pack kernel (peer stores over NVLink, CUDA IPC mapping),
        # twice: the second epoch exercises the ready / landed flag protocol on a reused buffer
        from bevy_firework_b200.distributed import PeerGather

        pg = PeerGather(eng, cap_rows_per_rank=max(counts) + 1024)
        for _ in range(2):
            pg.issue()
            prow, pcounts = pg.result()
        np.save(os.path.join(out_dir, f"prows{rank}.npy"), prow.cpu().numpy())
        assert pcounts == counts, (pcounts, counts)
        dist.barrier()
        pg.close()
        eng.close()
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_and_all_gather(tmp_path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "rows0.npy"), np.load(tmp_path / "rows1.npy")
    counts = np.load(tmp_path / "counts0.npy")
    assert r0.shape == r1.shape and (r0 == r1).all()      # both GPUs hold the whole scene
    for r in range(world):                                 # peer-store gather == NCCL gather, bit for bit
        assert np.load(tmp_path / f"prows{r}.npy").tobytes() == r0.tobytes()
    assert r0.shape[0] == counts.sum()
    # sharding does not change the result: a single GPU simulating all 8 spawners produces the
    # same rows (RNG streams are keyed by spawner, not by rank)
    from bevy_firework_b200._native import Engine

    eng = Engine(device=0, seed=0x00F12E00)
    _simulate(eng, range(N_SPAWNERS))
    buf = torch.empty((eng.total_live() + 16, 16), dtype=torch.float32, device="cuda:0")
    n = eng.pack_instances_device(buf.data_ptr(), buf.shape[0])
    single = buf[:n].cpu().numpy()
    eng.close()
    assert single.shape == r0.shape
    assert single.tobytes() == r0.tobytes()


def test_peer_gather_two_contexts_one_process():
    """one process driving two GPUs (how a single Bevy app would): plain peer access instead of IPC"""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from bevy_firework_b200._native import Engine
    from bevy_firework_b200.distributed import shard_range

    engs = [Engine(device=r, seed=0x00F12E00) for r in range(2)]
    for r, e in enumerate(engs):
        _simulate(e, shard_range(N_SPAWNERS, 2, r))
    lives = [e.total_live() for e in engs]
    handles = [e.gather_create(2, r, max(lives) + 256) for r, e in enumerate(engs)]
    for e in engs:
        e.gather_connect(handles)
    for e in engs:       # every rank issues before anyone waits
        e.gather_instances()
    got = []
    for e in engs:
        ptr, counts, stride = e.gather_result(2)
        assert counts == lives
        from bevy_firework_b200.distributed import _device_view

        dev = torch.device("cuda", e.device)
        got.append(torch.cat([_device_view(ptr + r * stride * 64, counts[r], dev) for r in range(2)]).cpu().numpy())
    assert got[0].tobytes() == got[1].tobytes()
    # against each context's own packed rows
    ref = []
    for e in engs:
        buf = torch.empty((e.total_live() + 16, 16), dtype=torch.float32, device=f"cuda:{e.device}")
        n = e.pack_instances_device(buf.data_ptr(), buf.shape[0])
        ref.append(buf[:n].cpu().numpy())
    assert np.concatenate(ref).tobytes() == got[0].tobytes()
    for e in engs:
        e.gather_destroy()
        e.close()
