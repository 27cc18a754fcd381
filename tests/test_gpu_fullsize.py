"""Row-for-row parity at the literal BASELINE.json sizes (configs[1..4] = C2, C3, C4, C5).

Every test replays the scene on the CUDA path (through the C ABI) and on the CPU oracle with the
same seed and dt = fl32(1/60), past the first deaths, and then compares EVERY field of EVERY row of
EVERY stream for equality -- no tolerance, no sampling (tests/_parity.py). Stream counts are
compared on every frame. The oracle runs one task per spawner on all host cores, as its timed
CPU-baseline form does; C5 alone uses its culling test helper (conservative boxes in front of the
brute-force ray loop, shown equal to it in tests/test_oracle_golden.py), because 1 M particles x 4
casts x 256 colliders x 130 frames of brute force would take minutes.
"""
import os

import numpy as np
import pytest

from bevy_firework_b200._native import frame_input
from bevy_firework_b200.workloads import (collision_ring, collision_scene_colliders, collision_spawner,
                                          grid_positions, one_shot_spawner, stress_spawner)
from _parity import assert_rows_match, reset_both

pytestmark = pytest.mark.gpu
DT = float(np.float32(1.0) / np.float32(60.0))
THREADS = os.cpu_count() or 8


def _oracle_counts(w, keys):
    return np.array([w._L.fwo_count(w._w, k, 0) for k in keys], dtype=np.uint64)


def _replay_grid(engine, oracle, n_spawners, rate, frames, min_live):
    w = oracle.OracleWorld(n_threads=THREADS)
    sp = stress_spawner(rate=rate)
    keys, inputs = [], []
    for i, p in enumerate(grid_positions(n_spawners)):
        reset_both(engine, w, 1 + i, sp)
        keys.append(1 + i)
        inputs.append(frame_input(1 + i, p))
    for k in range(frames):
        engine.frame(DT, inputs)
        w.frame(DT, inputs)
        gk, _, gc = engine.counts_all()
        assert (gk == np.array(keys, dtype=np.uint32)).all()
        assert (gc.astype(np.uint64) == _oracle_counts(w, keys)).all(), f"frame {k}"
    assert engine.total_live() == w.total_live() >= min_live
    for key in keys:
        assert_rows_match(engine.read_particles(key, 0), w.read_particles(key, 0), what=f"spawner {key}")
        bb, ob = engine.read_aabb(key), w.read_aabb(key)
        assert bb == ob, key
    w.close()


def test_c2_1m_particles_64_spawners_rows(engine, oracle):
    """configs[1]: stress_test.rs x 64 spawners at rate 15 625/s -> ~1 M live; 75 frames (lifetime 1 s:
    the first particles die on update #61), all 64 streams row for row"""
    _replay_grid(engine, oracle, 64, 15625.0, 75, 950_000)


def test_c3_10m_particles_512_spawners_rows(engine, oracle):
    """configs[2] on one GPU: stress_test.rs x 512 spawners at rate 19 531/s -> ~10 M live; 70 frames,
    all 512 streams row for row (the bench workload itself)"""
    _replay_grid(engine, oracle, 512, 19531.0, 70, 9_700_000)


def test_c4_literal_100k_bursts_rows(engine, oracle):
    """configs[3]: one new OneShot(100 000) spawner per frame with the one_shot.rs settings (lifetime
    2.5 s), retired when finished, 160 frames: ~15 M live at the end, bursts die on their update #151;
    every burst row for row at frame 80 and at the end, statuses on the way"""
    w = oracle.OracleWorld(n_threads=THREADS)
    sp = one_shot_spawner(100_000, 2.5)
    live = []
    for k in range(160):
        key = 1000 + k
        reset_both(engine, w, key, sp)
        live.append(key)
        a = 0.37 * k
        inp = [frame_input(key, (4.0 * np.cos(a), 1.0, 4.0 * np.sin(a)))]
        engine.frame(DT, inp)
        w.frame(DT, inp)
        assert engine.total_live() == w.total_live(), f"frame {k}"
        # notify_finished_particle_spawners (src/core.rs:674-688): retire what both sides call finished
        for old in list(live):
            ge, oe = engine.status(old), w.status(old)
            assert (ge.finished, ge.all_empty, ge.active) == (oe.finished, oe.all_empty, oe.active), (k, old)
            if not ge.finished:
                break
            engine.spawner_remove(old)
            w.spawner_remove(old)
            live.remove(old)
        if k in (80, 159):
            for key2 in live:
                assert_rows_match(engine.read_particles(key2, 0), w.read_particles(key2, 0), what=f"frame {k} burst {key2}")
    assert len(live) == 150 and engine.total_live() == 15_000_000
    w.close()


def test_c5_1m_particles_256_colliders_rows(engine, oracle):
    """configs[4]: stress_test_collision.rs x 8 spawners at rate 63 000/s vs 256 cuboids -> ~1 M live;
    130 frames (lifetime 2 s: deaths from update #121), every bounce of every particle bit-equal"""
    w = oracle.OracleWorld(n_threads=THREADS, cull=True)
    sp = collision_spawner(rate=63000.0)
    cols = collision_scene_colliders(256)
    engine.set_colliders(cols)
    w.set_colliders(cols)
    keys, inputs = [], []
    for i, (t, r) in enumerate(collision_ring(8)):
        reset_both(engine, w, 10 + i, sp)
        keys.append(10 + i)
        inputs.append(frame_input(10 + i, t, r))
    for k in range(130):
        engine.frame(DT, inputs)
        w.frame(DT, inputs)
        if k % 10 == 9:
            assert (engine.counts_all()[2].astype(np.uint64) == _oracle_counts(w, keys)).all(), f"frame {k}"
    assert engine.total_live() == w.total_live() >= 950_000
    for key in keys:
        assert_rows_match(engine.read_particles(key, 0), w.read_particles(key, 0), what=f"spawner {key}")
    w.close()
