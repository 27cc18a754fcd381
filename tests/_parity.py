"""Shared helpers of the parity tests: drive the CPU oracle and the CUDA path with the same
settings bytes and the same frame inputs, then compare ``ParticleData`` rows.

Bar: counts AND every field of every row are equal -- no tolerance anywhere (north_star allows
1e-5 relative; the build does not need it).
"""
from __future__ import annotations

import numpy as np

from bevy_firework_b200._native import frame_input

ALL_FIELDS = ("position", "velocity", "rotation", "angular_velocity", "initial_scale", "scale", "age",
              "lifetime", "base_color", "emissive_color", "pbr")


def same_bits(g, w):
    """IEEE equality (so -0 == +0) with NaNs required in the same places"""
    g = np.asarray(g)
    w = np.asarray(w)
    if g.dtype.kind != "f":
        return g == w
    return (g == w) | (np.isnan(g) & np.isnan(w))


def assert_rows_match(got, want, what=""):
    """every field of every ParticleData row equal, no tolerance: the kernels evaluate the
    reference's expressions in the reference's order with IEEE operations only (-fmad=false; sine and
    cosine from include/fw_sincos.h on both sides), so nothing may differ"""
    assert len(got) == len(want), f"{what}: count {len(got)} != oracle {len(want)}"
    for f in ALL_FIELDS:
        g, w = got[f], want[f]
        bad = ~same_bits(g, w)
        if bad.ndim > 1:
            bad = bad.any(axis=1)
        assert not bad.any(), (f"{what}: field {f} differs at row {int(np.argmax(bad))}: "
                               f"{g[np.argmax(bad)]} vs {w[np.argmax(bad)]} ({int(bad.sum())} of {len(g)} rows)")


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
