#!/usr/bin/env python
"""Render-extract all-gather on N GPUs (torchrun): the peer-store gather fused into the pack
kernel (fw_gather_*) against pack + NCCL all_gather. C3 per rank. Prints one JSON line."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bevy_firework_b200 import workloads as W  # noqa: E402
from bevy_firework_b200._native import Engine  # noqa: E402
from bevy_firework_b200.distributed import PeerGather, all_gather_instances  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    workload = sys.argv[1] if len(sys.argv) > 1 else "c3"
    reps = 10
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = Engine(device=local, seed=W.SEED)
    sc = bench.Scene(eng, workload, rank)
    for _ in range(sc.fill_frames + 5):
        sc.step()
    live = eng.total_live()
    pg = PeerGather(eng, cap_rows_per_rank=live + (1 << 16))

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def peer():
        pg.issue()
        eng.gather_result(world)

    def peer_with_frame():  # the usual use: one simulation frame, then the extract
        sc.step()
        pg.issue()
        eng.gather_result(world)

    def frame_only():
        sc.step()
        eng.sync()

    def nccl():
        all_gather_instances(eng)

    t_peer, t_nccl = timed(peer), timed(nccl)
    t_frame, t_both = timed(frame_only), timed(peer_with_frame)
    counts = eng.gather_result(world)[1]
    total = sum(counts)
    if rank == 0:
        print(json.dumps({
            "what": "all-gather-v of ParticleInstance rows (64 B) on every GPU", "workload": sc.label, "n_gpus": world,
            "rows_per_gpu": counts, "rows_total": total,
            "peer_store_gather_ms": t_peer * 1e3, "pack_plus_nccl_all_gather_ms": t_nccl * 1e3,
            "frame_ms": t_frame * 1e3, "frame_plus_peer_gather_ms": t_both * 1e3,
            "bytes_received_per_gpu_over_nvlink": (total - counts[0]) * 64,
            "peer_store_nvlink_gbs_per_gpu": (total - counts[0]) * 64 / t_peer / 1e9,
            "timing": "host wall clock around issue + fw_gather_result (sync), max over ranks, mean of %d" % reps}), flush=True)
    dist.barrier()
    pg.close()
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
