#!/usr/bin/env python
"""Writes tests/golden/*.npy: ParticleData rows of three small scenes after a fixed number of frames,
produced by the CPU ORACLE (the reference is Rust and cannot run here; its own two golden tests G1 / G2
are checked in tests/test_oracle_golden.py). The fixtures pin the oracle itself -- an accidental change
of the restatement, of include/fw_sincos.h or of the Philox protocol shows up as a diff against bytes
committed in an earlier round -- and give the `-m gpu` suite a comparison that does not depend on the
oracle library built at test time.   python scripts/gen_golden_fixtures.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bevy_firework_b200._native import frame_input  # noqa: E402
from bevy_firework_b200.workloads import (SEED, collision_ring, collision_scene_colliders, collision_spawner,  # noqa: E402
                                          sparks_spawner, stress_spawner)

DT = float(np.float32(1.0) / np.float32(60.0))


def scenes():
    """name -> (colliders, [(key, spawner, translation, rotation)], frames)"""
    ident = (0.0, 0.0, 0.0, 1.0)
    ring = collision_ring(2)
    return {
        "sparks_rate1000_90_frames": (None, [(1, sparks_spawner(1000.0), (0.0, 0.1, 0.0), ident)], 90),
        "stress_random_lifetime_2x_80_frames": (None, [(1, stress_spawner(rate=900.0, lifetime=0.6, lifetime_spread=0.4), (0.0, 0.1, 0.0), ident),
                                                       (2, stress_spawner(rate=700.0), (2.0, 0.1, 0.0), ident)], 80),
        "collision_2_spawners_64_colliders_100_frames": (collision_scene_colliders(64), [(10 + i, collision_spawner(rate=500.0), t, r)
                                                                                          for i, (t, r) in enumerate(ring)], 100),
    }


def run(backend, colliders, spawners, frames):
    if colliders:
        backend.set_colliders(colliders)
    inputs = []
    for key, sp, t, r in spawners:
        ps, nt, es, ne = sp.pods()
        backend.spawner_reset(key, ps, nt, es, ne, True)
        inputs.append(frame_input(key, t, r))
    for _ in range(frames):
        backend.frame(DT, inputs)
    return np.concatenate([backend.read_particles(key, 0) for key, *_ in spawners])


def main():
    from oracle import oracle as O

    out = os.path.join(ROOT, "tests", "golden")
    for name, (cols, spawners, frames) in scenes().items():
        w = O.OracleWorld(seed=SEED)
        rows = run(w, cols, spawners, frames)
        w.close()
        np.save(os.path.join(out, name + ".npy"), rows)
        print(name, len(rows), "rows", rows.nbytes, "bytes")


if __name__ == "__main__":
    main()
