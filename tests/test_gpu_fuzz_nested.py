"""A second seeded random campaign, over what tests/test_gpu_edge_and_scale.py::test_randomized_mixed_scene
leaves out: spawners with several particle types and Nested emitters (src/core.rs:471-546), the three
pacing kinds (:11-44) with cycles and windows, Local / Global spawn transforms (:66-73), spawners that
move, turn and change their modifiers every frame, queued particles, cylinder and cone colliders,
collision layers and SpatialQueryFilter exclusions (:240-248). Counts every frame, at the end every
field of every row of every stream equal to the oracle's."""
import os

import numpy as np
import pytest

from bevy_firework_b200 import (EmissionMode, EmissionPacing, EmissionSettings, EmissionShape, FireworkCurve,
                                FireworkGradient, LinearRgba, ParticleCollisionSettings, ParticleSettings,
                                ParticleSpawner, RandF32, RandVec3, SpawnTransformMode)
from bevy_firework_b200._native import frame_input
from bevy_firework_b200.workloads import capsule, cone, cuboid, cylinder, sphere
from _parity import assert_rows_match, reset_both

pytestmark = pytest.mark.gpu
f32 = np.float32


def _seeds():
    seeds = [1, 2, 3, 4]
    extra = os.environ.get("FW_FUZZ_SEEDS", "")
    if "-" in extra:
        a, b = extra.split("-")
        seeds += [s for s in range(int(a), int(b) + 1) if s not in seeds]
    return seeds


def _quat(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    return tuple(float(f32(c)) for c in q)


@pytest.mark.parametrize("seed", _seeds())
def test_randomized_nested_scene(engine, oracle, seed):
    rng = np.random.default_rng(7000 + seed)

    def col():
        return LinearRgba(*[float(v) for v in rng.uniform(0.0, 3.0, 4)])

    def gradient():
        k = rng.integers(0, 3)
        if k == 0:
            return FireworkGradient.constant(col())
        if k == 1:
            return FireworkGradient.even_samples([col() for _ in range(rng.integers(2, 5))])
        ts = np.sort(rng.uniform(0.05, 0.95, rng.integers(1, 4)))
        return FireworkGradient.uneven_samples([(0.0, col())] + [(float(t), col()) for t in ts] + [(1.0, col())])

    def curve():
        if rng.integers(0, 2):
            return FireworkCurve.constant(float(rng.uniform(0.5, 1.5)))
        return FireworkCurve.even_samples([float(v) for v in rng.uniform(0.2, 2.0, rng.integers(2, 5))])

    def pacing(nested):
        k = rng.integers(0, 4)
        if nested or k == 0:  # per-particle emitters count over the parent's age
            dur = float(rng.choice([0.0, 0.3, 1.0]))
            a = float(rng.uniform(0.0, 0.5))
            return EmissionPacing.CountOverDuration(float(rng.uniform(1.0, 12.0) if nested else rng.uniform(200.0, 5000.0)),
                                                    dur, a, float(rng.uniform(a, 1.0)))
        if k == 1:
            return EmissionPacing.OneShot(int(rng.integers(0, 3000)))
        if k == 2:
            return EmissionPacing.OnDemand
        return EmissionPacing.rate(float(rng.uniform(100.0, 8000.0)))

    def shape():
        return [EmissionShape.Point, EmissionShape.Sphere(float(rng.uniform(0.05, 0.5))),
                EmissionShape.Circle(tuple(float(c) for c in rng.normal(size=3)), float(rng.uniform(0.05, 0.5)))][rng.integers(0, 3)]

    def spawner():
        n_types = int(rng.integers(1, 4))
        types = []
        for _ in range(n_types):
            life = float(rng.uniform(0.15, 0.7))
            collision = None
            c = rng.integers(0, 4)
            if c:
                collision = ParticleCollisionSettings(float(rng.uniform(0.1, 0.9)), float(rng.uniform(0.0, 0.6)), c == 3,
                                                      filter=int(rng.choice([0xFFFFFFFF, 1, 2, 3])),
                                                      excluded=[int(k) for k in rng.choice(np.arange(100, 130), rng.integers(0, 4), replace=False)])
            types.append(ParticleSettings(
                lifetime=RandF32.constant(life) if rng.integers(0, 2) else RandF32(0.5 * life, 1.4 * life),
                initial_scale=RandF32(0.02, 0.15), scale_curve=curve(), base_color=gradient(), emissive_color=gradient(),
                acceleration=tuple(float(c) for c in rng.uniform(-6.0, 2.0, 3)),
                angular_acceleration=tuple(float(c) for c in rng.uniform(-2.0, 2.0, 3)) if rng.integers(0, 2) else (0.0, 0.0, 0.0),
                linear_drag=float(rng.uniform(0.0, 0.6)), angular_drag=float(rng.uniform(0.0, 0.6)),
                pbr=bool(rng.integers(0, 2)), collision_settings=collision))
        emitters = []
        for e in range(int(rng.integers(1, 5))):
            target = int(rng.integers(0, n_types))
            # parents of a lower type only: a type that feeds itself (covered with a small count in
            # test_gpu_nested_destroyed.py) or a cycle of types grows exponentially on both sides
            nested = e > 0 and target > 0 and rng.integers(0, 2) == 1
            emitters.append(EmissionSettings(
                particle_index=target, emission_pacing=pacing(nested),
                emission_mode=EmissionMode.Nested(int(rng.integers(0, target))) if nested else EmissionMode.Global,
                emission_shape=shape(),
                initial_velocity=RandVec3(RandF32(0.5, 6.0), tuple(float(c) for c in rng.normal(size=3)), float(rng.uniform(0.0, 1.2))),
                initial_velocity_radial=RandF32(0.0, float(rng.uniform(0.0, 1.5))),
                inherit_parent_velocity=bool(rng.integers(0, 2)), initial_rotation=_quat(rng),
                initial_angular_velocity=RandVec3(RandF32(0.0, 4.0), (0.0, 1.0, 0.0), 0.7) if rng.integers(0, 2)
                else RandVec3.constant((0.0, 0.0, 0.0))))
        return ParticleSpawner(particle_settings=types, emission_settings=emitters,
                               spawn_transform_mode=SpawnTransformMode.Local if rng.integers(0, 2) else SpawnTransformMode.Global)

    cols = [cuboid((30, 1, 30), (0, -0.5, 0), layers=3, key=100)]
    for i in range(30):
        p = (float(rng.uniform(-4, 4)), float(rng.uniform(0.3, 3.0)), float(rng.uniform(-4, 4)))
        layers = int(rng.choice([1, 2, 3]))
        k = rng.integers(0, 5)
        if k == 4:
            c = capsule(float(rng.uniform(0.2, 0.6)), float(rng.uniform(0.3, 1.5)), p, _quat(rng), layers=layers)
        elif k == 0:
            c = cuboid(tuple(rng.uniform(0.3, 1.5, 3)), p, _quat(rng), layers=layers, key=101 + i)
        elif k == 1:
            c = sphere(float(rng.uniform(0.2, 0.8)), p, layers=layers, key=101 + i)
        elif k == 2:
            c = cylinder(float(rng.uniform(0.2, 0.8)), float(rng.uniform(0.3, 1.5)), p, _quat(rng), layers=layers)
        else:
            c = cone(float(rng.uniform(0.2, 0.8)), float(rng.uniform(0.3, 1.5)), p, _quat(rng), layers=layers)
        cols.append(c)
    w = oracle.OracleWorld()
    engine.set_colliders(cols)
    w.set_colliders(cols)
    spawners = {}
    for key in range(1, 9):
        spawners[key] = spawner()
        reset_both(engine, w, key, spawners[key])
    pos = {key: rng.uniform(-3, 3, 3) + np.array([0.0, 4.0, 0.0]) for key in spawners}
    vel = {key: rng.uniform(-2, 2, 3) for key in spawners}
    for k in range(90):
        dt = float(f32(rng.choice([1 / 60, 1 / 60, 1 / 144, 1 / 30, 0.0])))
        if k == 40:
            key = int(rng.choice(list(spawners)))
            reset_both(engine, w, key, spawners[key])
        if k == 60:
            key = int(rng.choice(list(spawners)))
            engine.spawner_remove(key)
            w.spawner_remove(key)
            del spawners[key]
        inp = []
        for key in spawners:
            pos[key] = pos[key] + vel[key] * dt
            inp.append(frame_input(key, tuple(float(f32(c)) for c in pos[key]), _quat(rng) if k % 7 == 0 else (0.0, 0.0, 0.0, 1.0),
                                   tuple(float(f32(c)) for c in vel[key]), float(f32(rng.uniform(0.5, 1.5))),
                                   float(f32(rng.uniform(0.5, 1.5))), int(rng.integers(0, 40)) if rng.integers(0, 5) == 0 else 0))
        engine.frame(dt, inp)
        w.frame(dt, inp)
        for key, sp in spawners.items():
            assert engine.counts(key) == w.counts(key), f"frame {k} spawner {key}"
    for key, sp in spawners.items():
        for t in range(len(sp.particle_settings)):
            assert_rows_match(engine.read_particles(key, t), w.read_particles(key, t), what=f"spawner {key} type {t}")
