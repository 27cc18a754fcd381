#!/usr/bin/env python
"""bench.py -- particles updated/sec of the fused per-frame step (spawn + update) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c3r|c2|c1|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one frame of the hot path (fw_frame = plan + spawn + update kernels) over the whole
workload at fixed dt = fl32(1/60). Default workload: C3 = examples/stress_test.rs settings,
512 spawners x rate 19531 => ~10 M live particles PER GPU (1.6 GB of state, >> the 126 MB L2, so
no L2 flush is needed between steps). With N > 1 the 512 spawners are split over the ranks (BASELINE
config 3: shard by spawner, no data-path collective, strong scaling); the same run also measures every
rank carrying all 512 (weak) and reports it beside the headline.

One JSON line on rank 0 (see the contract in the task statement): value = whole-job particles/s
with state resident in HBM; e2e = same metric through the C ABI with host input structs and a
device->host read of the per-frame results (counts + AABBs) every step; roofline for the update
kernel from CUDA events recorded around it inside the timed region; cpu_baseline = the CPU oracle
(a C restatement of the reference loop; no Rust toolchain exists here) timed on the host cores.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from bevy_firework_b200 import workloads as W  # noqa: E402

ALGO_BYTES_PER_PARTICLE = 156  # 64 B read + 92 B written
# dominant kernel and its algorithmic bytes per particle, per workload: the
# compacting variant also moves the two constants (lifetime, initial_scale) with the particle
KERNEL_OF = {"c3r": ("fw::update_kernel<true,0>", 164), "c5": ("fw::update_kernel<false,1>", 156)}
DT = float(np.float32(1.0) / np.float32(60.0))


# ------------------------------------------------------------------------------ workloads
def make_workload(name: str, rank: int):
    """-> (label, [(key, ParticleSpawner, translation, rotation)], colliders, fill_frames, bursts)"""
    base = 1 + rank * 100000
    ident = (0.0, 0.0, 0.0, 1.0)
    if name == "c3":
        sp = W.stress_spawner(rate=19531.0)
        pos = W.grid_positions(512)
        return ("C3 stress_test.rs x512 spawners, rate 19531/s, lifetime 1 s (~10 M live particles per GPU)",
                [(base + i, sp, p, ident) for i, p in enumerate(pos)], None, 64, None)
    if name == "c3r":
        sp = W.stress_spawner(rate=19531.0, lifetime=1.0, lifetime_spread=0.5)
        pos = W.grid_positions(512)
        return ("C3r = C3 with lifetime U[0.5, 1.5] s: deaths anywhere in the Vec, in-place stable compaction "
                "(~10 M live particles per GPU)",
                [(base + i, sp, p, ident) for i, p in enumerate(pos)], None, 96, None)
    if name == "c3g":
        # C3's scene with NOTHING provable: angular velocity, emissive gradient, scale curve -> every
        # ParticleData field is per-particle state, the 156 B/particle layout of SURVEY section 8d
        from bevy_firework_b200 import FireworkCurve, FireworkGradient, LinearRgba, RandF32, RandVec3

        sp = W.stress_spawner(rate=19531.0)
        sp.particle_settings[0].scale_curve = FireworkCurve.even_samples([1.0, 0.5])
        sp.particle_settings[0].emissive_color = FireworkGradient.even_samples([LinearRgba(1.0, 0.5, 0.1, 1.0), LinearRgba(0.0, 0.0, 0.0, 0.0)])
        sp.emission_settings[0].initial_angular_velocity = RandVec3(RandF32(1.0, 4.0), (0.0, 1.0, 0.0), 0.5)
        pos = W.grid_positions(512)
        return ("C3g = C3 with angular velocity, an emissive gradient and a scale curve: every ParticleData field varies "
                "(generic 156 B/particle layout, ~10 M live particles per GPU)",
                [(base + i, sp, p, ident) for i, p in enumerate(pos)], None, 64, None)
    if name == "c2":
        sp = W.stress_spawner(rate=15625.0)
        pos = W.grid_positions(64)
        return ("C2 stress_test.rs x64 spawners, rate 15625/s, lifetime 1 s (~1 M live particles per GPU)",
                [(base + i, sp, p, ident) for i, p in enumerate(pos)], None, 64, None)
    if name == "c1":
        return ("C1 sparks.rs, 1 spawner, rate 6667/s, lifetime 0.75 s (~5 k live particles)",
                [(base, W.sparks_spawner(6667.0), (0.0, 0.1, 0.0), ident)], None, 50, None)
    if name == "c5":
        sp = W.collision_spawner(rate=63000.0)
        return ("C5 stress_test_collision.rs x8 spawners, rate 63000/s, lifetime 2 s, 256 cuboid colliders (~1 M live)",
                [(base + i, sp, t, r) for i, (t, r) in enumerate(W.collision_ring(8))],
                W.collision_scene_colliders(256), 124, None)
    if name == "c5d":
        # C5 with destroy_on_collision: a particle dies at its first hit, deaths are only known after the
        # sweep -> the compacting update with decoupled look-back (no BASELINE config; the variant's number)
        sp = W.collision_spawner(rate=125000.0)
        sp.particle_settings[0].collision_settings.destroy_on_collision = True
        return ("C5d = stress_test_collision.rs x8 spawners, rate 125000/s, destroy_on_collision, 256 cuboid colliders "
                "(look-back compaction; ~1 M live)",
                [(base + i, sp, t, r) for i, (t, r) in enumerate(W.collision_ring(8))],
                W.collision_scene_colliders(256), 124, None)
    if name == "c4":
        return ("C4 one_shot.rs: one new OneShot(100000) spawner per frame, lifetime 2.5 s (~15 M live)",
                [], None, 152, (W.one_shot_spawner(100000, 2.5), base))
    raise SystemExit(f"unknown workload {name}")


class Scene:
    """Drives either backend (Engine or OracleWorld: same method names) through the schedule."""

    def __init__(self, backend, workload, rank, shard=None, max_spawners=None):
        self.b = backend
        self.label, self.spawners, colliders, self.fill_frames, self.bursts = make_workload(workload, rank)
        if max_spawners is not None and len(self.spawners) > max_spawners:  # CPU arm: a bounded sample
            self.spawners = self.spawners[:max_spawners]
        if shard is not None:  # strong scaling: this rank simulates only its block of the spawners
            world, r = shard
            from bevy_firework_b200.distributed import shard_range

            base_keys = make_workload(workload, 0)[1]  # keys independent of the rank
            self.spawners = [base_keys[i] for i in shard_range(len(base_keys), world, r)]
        if colliders:
            backend.set_colliders(colliders)
        from bevy_firework_b200._native import frame_input

        self._frame_input = frame_input
        inputs = []
        for key, sp, t, r in self.spawners:
            ps, nt, es, ne = sp.pods()
            backend.spawner_reset(key, ps, nt, es, ne, True)
            inputs.append(frame_input(key, t, r))
        from bevy_firework_b200 import _abi

        self.inputs = (_abi.fw_spawner_frame_input * max(len(inputs), 1))(*inputs)
        self.n_inputs = len(inputs)
        self.frame_no = 0
        self.live_bursts = []
        self.updated = 0  # CPU arm: particles that entered the update so far

    def step(self):
        if self.bursts is not None:  # C4: new spawner each frame, retire the finished one
            sp, base = self.bursts
            key = base + self.frame_no
            ps, nt, es, ne = sp.pods()
            self.b.spawner_reset(key, ps, nt, es, ne, True)
            self.live_bursts.append(key)
            if len(self.live_bursts) > 151:  # lifetime 2.5 s: gone on update #151
                self.b.spawner_remove(self.live_bursts.pop(0))
            a = 0.37 * self.frame_no
            inputs = [self._frame_input(key, (4.0 * np.cos(a), 1.0, 4.0 * np.sin(a)))]
        else:
            inputs = None
        if isinstance(self.b, OracleBackendTag):
            # spawn_particles ; update_particles, counting the particles that enter the update
            self.b.spawn_only(DT, inputs if inputs is not None else self.inputs[: self.n_inputs])
            self.updated += self.b.total_live()
            self.b.update_only(DT)
        elif inputs is not None:
            self.b.frame(DT, inputs)
        else:
            self.b._check(self.b._L.fw_frame(self.b._ctx, DT, self.inputs, self.n_inputs))
        self.frame_no += 1


class OracleBackendTag:  # marker base so Scene knows which call convention to use
    pass


# ------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """NVML samples of SM clock and throttle reasons while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.01):
        super().__init__(daemon=True)
        self.period = period_s
        self.samples, self.reasons = [], set()
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max_mhz = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        return {"sm_mhz": (float(np.median(self.samples)) if self.samples else None),
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def ncu_traffic(workload: str):
    """per-launch DRAM traffic of the dominant kernel from the committed ncu capture (or None)"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)[workload]
        return t["dram_bytes_read"] + t["dram_bytes_write"], t["source"]
    except Exception:
        return None, None


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------ CPU arm
def run_cpu(workload: str, steps: int, warmup: int, threads: int, budget_s: float = 0.0):
    """the reference's CPU path: the C oracle (reference-faithful AoS loop, task per spawner).
    budget_s > 0: every step is a bounded sample of the workload -- the first spawners of it, as
    many as keep fill + warm-up + `steps` frames within about budget_s seconds (the loop costs
    ~1.8e-8 s per particle update on these hosts; throughput does not depend on the sample size
    as long as every thread has spawners)."""
    from oracle import oracle as O

    class Backend(O.OracleWorld, OracleBackendTag):
        pass

    b = Backend(seed=W.SEED, n_threads=threads)
    max_spawners = None
    if budget_s > 0.0:
        n_all = len(make_workload(workload, 0)[1])
        fill = make_workload(workload, 0)[3]
        if n_all:
            per_spawner_frame_s = {"c5": 8.0e-3, "c5d": 8.0e-3}.get(workload, 0.175 / 512.0 * (16.0 / max(threads, 1)))
            fit = int(budget_s / ((fill + warmup + steps) * per_spawner_frame_s))
            max_spawners = max(min(n_all, threads), min(n_all, fit - fit % max(threads, 1)))
    sc = Scene(b, workload, 0, max_spawners=max_spawners)
    for _ in range(sc.fill_frames + warmup):
        sc.step()
    sc.updated = 0
    t0 = time.perf_counter()
    for _ in range(steps):
        sc.step()
    dt = time.perf_counter() - t0
    updated = sc.updated
    live = b.total_live()
    b.close()
    sample = "full workload" if max_spawners is None or max_spawners >= len(make_workload(workload, 0)[1]) else \
        f"the first {max_spawners} spawners of the workload"
    return updated / dt, dt / steps * 1e3, live, sc.label, sample


# ------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c1", "c2", "c3", "c3g", "c3r", "c4", "c5", "c5d"])
    ap.add_argument("--cpu-steps", type=int, default=8, help="timed frames of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-budget-s", type=float, default=120.0,
                    help="--impl reference: bound of the whole CPU run in seconds (the sample shrinks for large --steps)")
    ap.add_argument("--no-extract", action="store_true", help="skip the e2e_extract leg (full instance-row extract, D2H)")
    ap.add_argument("--blocks", type=int, default=10, help="timed blocks of --steps frames each; value = the median block")
    ap.add_argument("--no-gather", action="store_true", help="N > 1: skip the render-extract gather leg")
    ap.add_argument("--single-scaling", action="store_true", help="N > 1: do not also measure the other scaling mode")
    ap.add_argument("--no-graphs", action="store_true", help="launch kernel by kernel instead of replaying frame graphs")
    ap.add_argument("--no-concurrent-spawn", action="store_true", help="run the spawn kernel before the update kernel")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong (default): the workload's spawners are sharded over the ranks (BASELINE config 3: the "
                         "10 M-particle scene over 1..8 GPUs); weak: every rank runs the whole workload. With N > 1 the "
                         "other mode is measured too and reported beside the headline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1

    # -------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        v, ms, live, label, sample = run_cpu(args.workload, args.steps, args.warmup, cores, budget_s=args.ref_budget_s)
        line = {
            "impl": "reference", "metric": "particles updated/sec (fused step)", "value": v, "unit": "particles/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": label, "dt": "fl32(1/60)", "live_particles": live,
                       "note": "one CPU workload on rank 0 whatever --gpus is"},
            "cpu_baseline": {"value": v, "unit": "particles/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} ({live} live particles), {args.steps} frames after {args.warmup} warm-up frames; "
                                       "C restatement of the reference loop (no Rust toolchain in the image), "
                                       f"one task per spawner on {cores} threads, sequential spawn"},
            "e2e": {"value": v, "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
        return 0

    # -------------------------------------------------------------- our arm (GPU)
    import torch
    import torch.distributed as dist

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    from bevy_firework_b200._native import Engine

    def make_engine():
        return Engine(device=local_rank, seed=W.SEED, profile=False, graphs=not args.no_graphs,
                      concurrent_spawn=not args.no_concurrent_spawn)  # raises without the CUDA library

    def barrier(e):
        e.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def reduce_max(vals):
        if world == 1:
            return list(vals)
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def reduce_sum(vals):
        if world == 1:
            return [int(v) for v in vals]
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [int(v) for v in t.tolist()]

    def gather_ranks(v):
        if world == 1:
            return [v]
        out = [None] * world
        dist.all_gather_object(out, v)
        return out

    sampler = ClockSampler(local_rank, period_s=0.02)
    sampler.start()  # before the warm-up: the timed region of a 20-step run is a few milliseconds

    def timed_blocks(e, scene, n_blocks):
        """n_blocks x [barrier + synchronize | EXACTLY `steps` frames between two CUDA events on the
        launching stream | barrier + synchronize]; per block: max over ranks of the time, sum over
        ranks of the particles. -> (block ms (max over ranks), particles per block, this rank's ms)"""
        out_ms, out_upd, mine = [], [], []
        for _ in range(n_blocks):
            barrier(e)
            e.profile_reset()
            e.event_record(0)
            for _ in range(args.steps):
                scene.step()
            e.event_record(1)
            ms_b = e.event_elapsed_ms(0, 1)
            barrier(e)
            prof_b, n_b = e.profile_sum()
            assert n_b == args.steps, (n_b, args.steps)
            out_ms.append(ms_b)
            out_upd.append(int(prof_b.particles_updated))
            mine.append(ms_b)
        launches_b = int(e.profile_sum()[0].kernel_launches)
        ms_all_b = reduce_max(out_ms)
        upd_all_b = reduce_sum(out_upd)
        return ms_all_b, upd_all_b, mine, launches_b

    def run_scene(scaling):
        e = make_engine()
        scene = Scene(e, args.workload, rank, shard=(world, rank) if scaling == "strong" and world > 1 else None)
        for _ in range(scene.fill_frames):  # reach the stationary live count (lifetime/dt + 2 frames)
            scene.step()
        e.sync()
        for _ in range(args.warmup):
            scene.step()
        ms_b, upd_b, mine, launches_b = timed_blocks(e, scene, args.blocks)
        k = int(np.argsort(ms_b)[len(ms_b) // 2])  # the median block
        return e, scene, {"ms": ms_b[k], "updated": upd_b[k], "block_ms_per_step": [m / args.steps for m in ms_b],
                          "rank_ms_per_step": gather_ranks(float(np.median(mine)) / args.steps), "launches": launches_b}

    # ---- device-resident throughput. N > 1: BASELINE config 3 is the 10 M / 512-spawner scene SPLIT
    # over the GPUs (spawner i -> GPU i // (512 / N)), so that is the headline ("strong"); the same
    # run also measures every GPU carrying the whole scene ("weak") and reports it beside it.
    eng, sc, main_t = run_scene(args.scaling)
    live = eng.total_live()
    # what one update moves per particle of this workload's streams (fw_stream_layout_get: fields the
    # library proved constant for the stream are not stored), and the kernel instantiation that runs
    first_key = sc.spawners[0][0] if sc.spawners else sc.live_bursts[-1]
    lay = eng.stream_layout(first_key, 0)
    layout_bytes = int(lay.bytes_read + lay.bytes_written)
    # (static streams without a collision sweep run the segment-scheduled kernel, everything else the tile-scheduled one)
    layout_kernel = ("fw::update_static_kernel<%s>" % ("true" if lay.variant & 1 else "false") if not lay.variant & 6 else
                     "fw::update_kernel<%s,%d,%s>" % ("true" if lay.variant & 1 else "false", 1 if lay.variant & 2 else 0,
                                                      "true" if lay.variant & 4 else "false"))
    layout_info = {"variant": int(lay.variant), "flags": int(lay.flags), "bytes_read": int(lay.bytes_read),
                   "bytes_written": int(lay.bytes_written), "bytes_count_pass": int(lay.bytes_count_pass),
                   "generic_bytes_per_particle": ALGO_BYTES_PER_PARTICLE}

    # ---- per-kernel durations: K more steps launched kernel by kernel with CUDA events around
    # every kernel (events recorded inside a replayed graph carry no timestamps)
    barrier(eng)
    eng.profile_reset()
    eng.set_profiling(True)
    for _ in range(args.steps):
        sc.step()
    barrier(eng)
    kprof, kn = eng.profile_sum()
    eng.set_profiling(False)
    assert kprof.timed_frames == args.steps, (kprof.timed_frames, args.steps)

    # ---- end to end through the C ABI: host inputs in, per-frame results (counts, AABBs) out
    keys = [k for k, *_ in sc.spawners]
    eng.profile_reset()
    barrier(eng)
    t0 = time.perf_counter()
    eng.event_record(2)
    e2e_steps = max(10, min(args.steps, 200))
    for _ in range(e2e_steps):
        sc.step()
        k_, t_, counts = eng.counts_all()      # waits for the frame's state readback (D2H), exact counts
        if keys:
            eng.read_aabb(keys[0])             # served from the same snapshot
    eng.event_record(3)
    ms_e2e_dev = eng.event_elapsed_ms(2, 3)
    ms_e2e_wall = (time.perf_counter() - t0) * 1e3
    prof2, n2 = eng.profile_sum()
    ms_e2e = max(ms_e2e_dev, ms_e2e_wall)
    updated_e2e = int(prof2.particles_updated)
    h2d = int(prof2.h2d_bytes // max(n2, 1))
    d2h = int(prof2.d2h_bytes // max(n2, 1))  # the frame's own state readback serves fw_counts_all / fw_read_aabb
    ms_e2e_all = reduce_max([ms_e2e])[0]
    updated_e2e_all = reduce_sum([updated_e2e])[0]

    # ---- end to end INCLUDING the render hand-off: every frame also brings the 64-byte
    # ParticleInstance rows of all live particles (what extract_firework_components consumes,
    # reference src/render.rs:439-461) into pinned host memory. PCIe-bound.
    extract = None
    if not args.no_extract:
        x_steps = 10
        cap = eng.total_live() + 4 * max(len(sc.spawners), 151) * 2048
        host = torch.empty((cap, 16), dtype=torch.float32, pin_memory=True)
        eng.extract_instances(host.data_ptr(), cap)
        barrier(eng)
        upd0 = eng.profile_sum()[0].particles_updated
        t0 = time.perf_counter()
        rows = 0
        for _ in range(x_steps):
            sc.step()
            rows = eng.extract_instances(host.data_ptr(), cap)
        dt_x = time.perf_counter() - t0
        upd1 = eng.profile_sum()[0].particles_updated
        dt_x_all = reduce_max([dt_x])[0]
        upd_x_all = reduce_sum([upd1 - upd0])[0]
        extract = {"value": upd_x_all / dt_x_all, "unit": "particles/s", "d2h_bytes_per_step": rows * 64,
                   "ms_per_step": dt_x_all / x_steps * 1e3, "steps": x_steps,
                   "pcie_gbs": rows * 64 / (dt_x / x_steps) / 1e9,
                   "what": "fw_frame + fw_extract_instances: all live 64-byte ParticleInstance rows into pinned host memory every step"}
        if world == 1:
            # the same hand-off, asynchronous and double-buffered (fw_extract_begin / fw_extract_wait: the copy
            # of frame f runs under frame f+1), and for a culled eighth of the spawners; against the plain
            # pinned-memory D2H rate of this box for the same bytes
            host2 = torch.empty((cap, 16), dtype=torch.float32, pin_memory=True)
            dev = torch.empty((max(rows, 1), 16), dtype=torch.float32, device="cuda")
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 0.0
            for _ in range(3):
                ev0.record()
                host[: dev.shape[0]].copy_(dev, non_blocking=True)
                ev1.record()
                ev1.synchronize()
                best = max(best, dev.numel() * 4 / (ev0.elapsed_time(ev1) * 1e-3) / 1e9)
            del dev
            bufs = [host, host2]
            barrier(eng)
            sc.step()
            eng.extract_begin(bufs[0].data_ptr(), cap)
            t0 = time.perf_counter()
            for i in range(x_steps):
                sc.step()
                eng.extract_begin(bufs[(i + 1) & 1].data_ptr(), cap)
                rows_a, _ = eng.extract_wait(0)
            dt_a = time.perf_counter() - t0
            eng.extract_wait(0)
            keys_all = [k for k, *_ in sc.spawners]
            sub = keys_all[:: 8] if keys_all else None
            culled = None
            if sub:
                barrier(eng)
                t0 = time.perf_counter()
                for i in range(x_steps):
                    sc.step()
                    eng.extract_begin(bufs[i & 1].data_ptr(), cap, sub)
                    rows_c, _ = eng.extract_wait(0)
                dt_c = time.perf_counter() - t0
                culled = {"spawners": len(sub), "rows": rows_c, "ms_per_step": dt_c / x_steps * 1e3}
            extract["pinned_d2h_peak_gbs"] = best
            extract["frac_of_pinned_d2h_peak"] = extract["pcie_gbs"] / best if best else None
            extract["async_double_buffered"] = {"ms_per_step": dt_a / x_steps * 1e3, "pcie_gbs": rows_a * 64 / (dt_a / x_steps) / 1e9,
                                                "frac_of_pinned_d2h_peak": rows_a * 64 / (dt_a / x_steps) / 1e9 / best if best else None,
                                                "what": "fw_extract_begin(frame f+1) issued before fw_extract_wait(frame f)"}
            extract["culled_subset"] = culled
            del host2
        del host

    live_all, launches_all = reduce_sum([live, main_t["launches"]])
    clocks_now = sampler.stop()

    # ---- N > 1: the one exchange of the path, the render extract -- every GPU's ParticleInstance rows
    # on every GPU (reference consumer src/render.rs:439-461). fw_gather_instances (pack kernel storing
    # into every rank's buffer over NVLink) against pack + NCCL all-gather; rows compared for equality.
    gather = None
    if world > 1 and not args.no_gather:
        from bevy_firework_b200.distributed import PeerGather, all_gather_instances

        rows_nccl, counts = all_gather_instances(eng)
        pg = PeerGather(eng, cap_rows_per_rank=max(counts) + 4096)
        pg.issue()
        rows_peer, pcounts = pg.result()
        same = bool(pcounts == counts and rows_peer.shape == rows_nccl.shape and torch.equal(rows_peer, rows_nccl))
        del rows_peer
        reps = 5
        barrier(eng)
        t0 = time.perf_counter()
        for _ in range(reps):
            pg.issue()
            eng.gather_result(world)  # waits for this rank's landed flags
        ms_peer = (time.perf_counter() - t0) / reps * 1e3
        barrier(eng)
        t0 = time.perf_counter()
        for _ in range(reps):
            r_, c_ = all_gather_instances(eng)
            torch.cuda.synchronize()
        ms_nccl = (time.perf_counter() - t0) / reps * 1e3
        del r_, rows_nccl
        ms_peer_all, ms_nccl_all = reduce_max([ms_peer, ms_nccl])
        same_all = min(gather_ranks(same))
        recv = (sum(counts) - min(counts)) * 64  # bytes the busiest link receives
        gather = {"rows_total": int(sum(counts)), "peer_store_ms": ms_peer_all, "nccl_ms": ms_nccl_all,
                  "received_gbs_per_gpu": recv / (ms_peer_all * 1e-3) / 1e9, "peer_copy_peak_gbs": 770.0,
                  "frac_of_peer_copy_peak": recv / (ms_peer_all * 1e-3) / 1e9 / 770.0,
                  "rows_equal_to_nccl_path": bool(same_all), "reps": reps,
                  "what": "fw_gather_instances + fw_gather_result (pack kernel with peer stores, device-side flags) vs "
                          "fw_pack_instances_device + NCCL all_gather; wall clock incl. the host sync, max over ranks"}
        pg.close()

    # ---- weak scaling beside the strong headline (N > 1), same run
    other = None
    if world > 1 and not args.single_scaling:
        eng.close()
        other_kind = "weak" if args.scaling == "strong" else "strong"
        eng2, sc2, ot = run_scene(other_kind)
        live2 = reduce_sum([eng2.total_live()])[0]
        other = {"scaling": other_kind, "value": ot["updated"] / (ot["ms"] * 1e-3), "unit": "particles/s",
                 "ms_per_step": ot["ms"] / args.steps, "live_particles": live2,
                 "block_ms_per_step": ot["block_ms_per_step"], "rank_ms_per_step": ot["rank_ms_per_step"]}
        eng2.close()
        eng = None

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        upd_kernel_ms = kprof.update_ms / args.steps
        per_launch_particles = int(kprof.particles_updated) / args.steps
        kernel_name, algo_bytes = layout_kernel, layout_bytes
        achieved = algo_bytes * per_launch_particles / (upd_kernel_ms * 1e-3) / 1e9
        ms_all, updated_all = main_t["ms"], main_t["updated"]
        line = {
            "metric": "particles updated/sec (fused step)", "value": updated_all / (ms_all * 1e-3), "unit": "particles/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_all / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": sc.label, "dt": "fl32(1/60)", "live_particles": live_all, "seed": hex(W.SEED),
                       "streams_per_gpu": len(sc.spawners) or 151,
                       "parallelism": (f"shard-by-spawner x{world}: the workload's spawners split over the GPUs" if args.scaling == "strong"
                                       else f"shard-by-spawner x{world}: every GPU carries the whole workload"),
                       "l2": "state per GPU (0.8 GB at C3 on one GPU) is larger than the 126 MB L2; no flush between steps"
                             + ("; split over 8 GPUs the 100 MB per GPU fit in L2 -- said here, not hidden" if world > 1 and args.scaling == "strong" else ""),
                       "fill_frames": sc.fill_frames, "cuda_graphs": not args.no_graphs,
                       "concurrent_spawn": not args.no_concurrent_spawn,
                       "timing": f"{args.blocks} blocks of exactly {args.steps} steps, each bracketed by barrier + synchronize and timed with "
                                 "CUDA events on the launching stream, max over ranks per block; value = the median block"},
            "block_ms_per_step": main_t["block_ms_per_step"], "rank_ms_per_step": main_t["rank_ms_per_step"],
            "e2e": {"value": updated_e2e_all / (ms_e2e_all * 1e-3), "unit": "particles/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "what": "fw_frame with host input structs + fw_counts_all/fw_read_aabb (sync + D2H) every step"},
            "gpu_launches": int(launches_all),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(args.workload)[0], "traffic_source": ncu_traffic(args.workload)[1],
                         "kernel": kernel_name,
                         "algorithmic_bytes_per_launch": algo_bytes * per_launch_particles,
                         "algorithmic_bytes_per_particle": algo_bytes,
                         "particles_per_launch": per_launch_particles, "kernel_ms": upd_kernel_ms, "layout": layout_info,
                         "peak_source": peak_src,
                         "generic_layout": {"algorithmic_bytes_per_particle": ALGO_BYTES_PER_PARTICLE,
                                            "note": "SURVEY 8d's figure for a stream where every ParticleData field varies; this "
                                                    "workload's streams keep only the fields that can vary (fw_stream_layout_get)"},
                         "how": f"CUDA events around the kernel, mean of {args.steps} launches in a second timed "
                                "region of the same run (kernel-by-kernel launches)"},
            "kernel_ms": {"plan": kprof.plan_ms / args.steps, "spawn": kprof.spawn_ms / args.steps,
                          "update": upd_kernel_ms, "frame": kprof.total_ms / args.steps},
            "clocks": clocks_now,
        }
        if extract:
            line["e2e_extract"] = extract
        if other:
            line[other["scaling"]] = other
        if gather:
            line["render_extract_gather"] = gather
        if world > 1:
            line["reference_arm_note"] = "the --impl reference arm times ONE CPU workload on rank 0 whatever N is"
        if world == 1 and not args.no_cpu_baseline:
            if eng is not None:
                eng.close()
                eng = None
            line["parity_checked"] = parity_check(args.workload, local_rank)
            v, cms, clive, _, _ = run_cpu(args.workload, args.cpu_steps, 1, cores)
            # Bevy's default compute pool does not get every core: 4-thread figure too
            v4 = run_cpu(args.workload, max(2, args.cpu_steps // 2), 1, 4)[0] if cores > 4 else v
            line["cpu_baseline"] = {
                "value": v, "unit": "particles/s", "cores": cores, "kind": "port", "value_4_threads": v4,
                "sample": f"full workload ({clive} live particles), {args.cpu_steps} timed frames after the fill "
                          f"frames, {cms:.1f} ms/frame; C restatement of the reference loop (oracle/fw_oracle.c), "
                          f"one task per spawner on {cores} threads, sequential spawn"}
        print(json.dumps(line), flush=True)
    if eng is not None:
        eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def parity_check(workload: str, device: int):
    """the bench workload's first spawners replayed on the CUDA path and on the CPU oracle, every
    field of every row compared for equality (the tests do this for the whole workload)"""
    from bevy_firework_b200._native import Engine
    from oracle import oracle as O

    class Backend(O.OracleWorld, OracleBackendTag):
        pass

    n = {"c5": 2, "c5d": 2, "c4": 0}.get(workload, 4)
    if n == 0:
        return None
    e = Engine(device=device, seed=W.SEED)
    b = Backend(seed=W.SEED, n_threads=4, cull=workload in ("c5", "c5d"))
    se, sb = Scene(e, workload, 0, max_spawners=n), Scene(b, workload, 0, max_spawners=n)
    frames = {"c5": 130, "c5d": 130}.get(workload, 70)
    for _ in range(frames):
        se.step()
        sb.step()
    rows = 0
    ok = True
    for key, *_ in se.spawners:
        g, w = e.read_particles(key, 0), b.read_particles(key, 0)
        ok = ok and len(g) == len(w) and all((g[f] == w[f]).all() for f in g.dtype.names)
        rows += len(g)
    e.close()
    b.close()
    if not ok:
        raise SystemExit("bench.py: the CUDA path and the CPU oracle disagree on the bench workload")
    return {"spawners": n, "frames": frames, "rows": rows, "equal": True}


if __name__ == "__main__":
    sys.exit(main())
