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
