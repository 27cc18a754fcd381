"""Shared helpers of the parity tests: drive the CPU oracle and the CUDA path with the same
settings bytes and the same frame inputs, then compare ``ParticleData`` rows.

Tolerances: counts are exact; every field whose value does not pass
through sinf/cosf is compared bit-for-bit (``exact`` list); the rest must satisfy
``abs(a-b) <= tol * max(abs(a), abs(b), 1)`` with tol = 1e-5 (north_star's bound).
"""
from __future__ import annotations

import numpy as np

from bevy_firework_b200._native import frame_input

TOL = 1e-5
ALL_FIELDS = ("position", "velocity", "rotation", "angular_velocity", "initial_scale", "scale", "age",
              "lifetime", "base_color", "emissive_color", "pbr")


def close(a, b, tol=TOL):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) <= tol * np.maximum(np.maximum(np.abs(a), np.abs(b)), 1.0)


def assert_rows_match(got, want, exact=(), tol=TOL, what=""):
    assert len(got) == len(want), f"{what}: count {len(got)} != oracle {len(want)}"
    for f in ALL_FIELDS:
        g, w = got[f], want[f]
        if f in exact or f == "pbr":
            bad = ~(g == w)
            if bad.ndim > 1:
                bad = bad.any(axis=1)
            assert not bad.any(), (f"{what}: field {f} not bit-exact at {int(np.argmax(bad))}: "
                                   f"{g[np.argmax(bad)]} vs {w[np.argmax(bad)]} ({int(bad.sum())} rows)")
        else:
            ok = close(g, w, tol)
            if ok.ndim > 1:
                ok = ok.all(axis=1)
            assert ok.all(), (f"{what}: field {f} off at {int(np.argmin(ok))}: "
                              f"{g[np.argmin(ok)]} vs {w[np.argmin(ok)]} ({int((~ok).sum())} rows)")


def reset_both(engine, world, key, spawner):
    ps, n_t, es, n_e = spawner.pods()
    engine.spawner_reset(key, ps, n_t, es, n_e, spawner.starts_enabled)
    world.spawner_reset(key, ps, n_t, es, n_e, spawner.starts_enabled)


def random_rows(rng, n, lifetime=(0.5, 3.0), angular=True):
    from bevy_firework_b200 import _abi

    rows = np.zeros(n, dtype=_abi.particle_data_dtype())
    rows["position"] = rng.uniform(-5, 5, (n, 3))
    rows["velocity"] = rng.uniform(-10, 10, (n, 3))
    q = rng.normal(size=(n, 4))
    rows["rotation"] = q / np.linalg.norm(q, axis=1, keepdims=True)
    if angular:
        rows["angular_velocity"] = rng.uniform(-6, 6, (n, 3))
    rows["initial_scale"] = rng.uniform(0.02, 0.5, n)
    rows["scale"] = rows["initial_scale"]
    rows["lifetime"] = rng.uniform(lifetime[0], lifetime[1], n)
    rows["age"] = rows["lifetime"] * rng.uniform(0.0, 0.98, n).astype(np.float32)
    rows["base_color"] = rng.uniform(0, 1, (n, 4))
    rows["emissive_color"] = rng.uniform(0, 1, (n, 4))
    return rows
