"""The reference's example scenes as workload definitions (BASELINE.json
``configs``). Each function returns ``ParticleSpawner`` settings copied from the cited example;
the scaling (how many spawners, which rate) is what BASELINE.json asks for.
"""
from __future__ import annotations

import math
from typing import List, Tuple

from . import _abi
from .core import (EmissionPacing, EmissionSettings, ParticleCollisionSettings, ParticleSettings,
                   ParticleSpawner, RandF32, RandVec3, SpawnTransformMode)
from .curve import FireworkCurve, FireworkGradient, LinearRgba
from .emission_shape import EmissionShape

DT_60HZ = 1.0 / 60.0  # converted to f32 at the ABI: fl32(1/60)
SEED = 0x00F12E00


def _fire_gradient(first: Tuple[float, float, float, float]) -> FireworkGradient:
    # examples/stress_test.rs:100-106, sparks.rs:57-63, stress_test_collision.rs:101-107
    return FireworkGradient.uneven_samples([
        (0.0, LinearRgba(*first)),
        (0.7, LinearRgba(3.0, 1.0, 1.0, 1.0)),
        (0.8, LinearRgba(1.0, 0.3, 0.3, 1.0)),
        (0.9, LinearRgba(0.3, 0.3, 0.3, 1.0)),
        (1.0, LinearRgba(0.1, 0.1, 0.1, 0.0)),
    ])


def sparks_spawner(rate: float = 1000.0) -> ParticleSpawner:
    """examples/sparks.rs:49-84 (C1). ``rate=6667`` gives the ~5 k live particles BASELINE.json
    quotes; the literal example uses 1000."""
    return ParticleSpawner(
        particle_settings=[ParticleSettings(
            lifetime=RandF32.constant(0.75),
            initial_scale=RandF32(0.02, 0.08),
            scale_curve=FireworkCurve.constant(1.0),
            base_color=_fire_gradient((150.0, 100.0, 15.0, 1.0)),
            linear_drag=0.1,
            pbr=False,
        )],
        emission_settings=[EmissionSettings(
            emission_pacing=EmissionPacing.rate(rate),
            emission_shape=EmissionShape.Circle((0.0, 1.0, 0.0), 0.3),
            inherit_parent_velocity=True,
            initial_velocity=RandVec3(RandF32(0.0, 10.0), (0.0, 1.0, 0.0), 30.0 / 180.0 * math.pi),
        )],
    )


def stress_spawner(rate: float = 160000.0, lifetime: float = 1.0, lifetime_spread: float = 0.0) -> ParticleSpawner:
    """examples/stress_test.rs:91-129 (C2/C3). ``lifetime_spread`` > 0 draws the lifetime from
    [lifetime - spread, lifetime + spread] (deaths anywhere in the Vec: the compacting update)."""
    return ParticleSpawner(
        particle_settings=[ParticleSettings(
            lifetime=RandF32(lifetime - lifetime_spread, lifetime + lifetime_spread) if lifetime_spread else RandF32.constant(lifetime),
            initial_scale=RandF32(0.02, 0.08),
            scale_curve=FireworkCurve.constant(1.0),
            base_color=_fire_gradient((10.0, 7.0, 1.0, 1.0)),
            linear_drag=0.1,
            pbr=False,
        )],
        emission_settings=[EmissionSettings(
            emission_pacing=EmissionPacing.rate(rate),
            emission_shape=EmissionShape.Circle((0.0, 1.0, 0.0), 0.3),
            inherit_parent_velocity=True,
            initial_velocity=RandVec3(RandF32(0.0, 10.0), (0.0, 1.0, 0.0), 30.0 / 180.0 * math.pi),
        )],
    )


def one_shot_spawner(count: int = 100_000, lifetime: float = 2.5) -> ParticleSpawner:
    """examples/one_shot.rs:92-130 (C4); initial_scale fixed to [0.1, 0.3] (the example derives
    it from a collision impulse, :96-99)."""
    return ParticleSpawner(
        particle_settings=[ParticleSettings(
            lifetime=RandF32.constant(lifetime),
            initial_scale=RandF32(0.1, 0.3),
            scale_curve=FireworkCurve.even_samples([1.0, 2.0]),
            base_color=FireworkGradient.uneven_samples([
                (0.0, LinearRgba(0.6, 0.3, 0.0, 0.0)),
                (0.1, LinearRgba(0.6, 0.3, 0.0, 0.35)),
                (1.0, LinearRgba(0.6, 0.3, 0.0, 0.0)),
            ]),
            linear_drag=0.7,
            pbr=True,
            acceleration=(0.0, -1.5, 0.0),
            fade_scene=3.5,
        )],
        emission_settings=[EmissionSettings(
            emission_pacing=EmissionPacing.OneShot(count),
            emission_shape=EmissionShape.Circle((0.0, 1.0, 0.0), 0.4),
            inherit_parent_velocity=True,
            initial_velocity=RandVec3(RandF32(0.0, 2.0), (0.0, 1.0, 0.0), 0.0),
            initial_velocity_radial=RandF32(0.0, 2.5),
        )],
        spawn_transform_mode=SpawnTransformMode.Local,
    )


def collision_spawner(rate: float = 80000.0) -> ParticleSpawner:
    """examples/stress_test_collision.rs:91-139 (C5)."""
    return ParticleSpawner(
        particle_settings=[ParticleSettings(
            lifetime=RandF32.constant(2.0),
            initial_scale=RandF32(0.02, 0.08),
            scale_curve=FireworkCurve.constant(1.0),
            linear_drag=0.15,
            base_color=_fire_gradient((100.0, 70.0, 10.0, 1.0)),
            pbr=False,
            collision_settings=ParticleCollisionSettings(restitution=0.6, friction=0.2,
                                                         destroy_on_collision=False),
        )],
        emission_settings=[EmissionSettings(
            emission_pacing=EmissionPacing.rate(rate),
            emission_shape=EmissionShape.Circle((0.0, 1.0, 0.0), 0.3),
            initial_velocity=RandVec3(RandF32(6.0, 8.0), (0.0, 1.0, 0.0), 30.0 / 180.0 * math.pi),
            inherit_parent_velocity=True,
        )],
    )


def quat_from_rotation_z(angle: float):
    return (0.0, 0.0, math.sin(angle * 0.5), math.cos(angle * 0.5))


def quat_from_rotation_x(angle: float):
    return (math.sin(angle * 0.5), 0.0, 0.0, math.cos(angle * 0.5))


def quat_from_rotation_y(angle: float):
    return (0.0, math.sin(angle * 0.5), 0.0, math.cos(angle * 0.5))


def quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return (aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz)


def cuboid(size, translation, rotation=(0.0, 0.0, 0.0, 1.0), layers: int = 1, key: int = _abi.FW_NO_KEY) -> _abi.fw_collider:
    """``Collider::cuboid(x, y, z)`` takes full extents (examples/stress_test_collision.rs:88)."""
    c = _abi.fw_collider()
    c.kind = _abi.FW_COLLIDER_CUBOID
    c.layers = layers
    c.key = key
    c.half_extents[:] = [0.5 * float(s) for s in size]
    c.translation[:] = [float(t) for t in translation]
    c.rotation[:] = [float(r) for r in rotation]
    return c


def sphere(radius, translation, layers: int = 1, key: int = _abi.FW_NO_KEY) -> _abi.fw_collider:
    c = _abi.fw_collider()
    c.kind = _abi.FW_COLLIDER_SPHERE
    c.layers = layers
    c.key = key
    c.half_extents[:] = [float(radius), 0.0, 0.0]
    c.translation[:] = [float(t) for t in translation]
    c.rotation[:] = [0.0, 0.0, 0.0, 1.0]
    return c


def _revolved(kind, radius, height, translation, rotation, layers) -> _abi.fw_collider:
    c = _abi.fw_collider()
    c.kind = kind
    c.layers = layers
    c.key = _abi.FW_NO_KEY
    c.half_extents[:] = [float(radius), 0.5 * float(height), 0.0]
    c.translation[:] = [float(t) for t in translation]
    c.rotation[:] = [float(r) for r in rotation]
    return c


def cylinder(radius, height, translation, rotation=(0.0, 0.0, 0.0, 1.0), layers: int = 1) -> _abi.fw_collider:
    """``Collider::cylinder(radius, height)``, axis +Y (examples/textures.rs:195)."""
    return _revolved(_abi.FW_COLLIDER_CYLINDER, radius, height, translation, rotation, layers)


def cone(radius, height, translation, rotation=(0.0, 0.0, 0.0, 1.0), layers: int = 1) -> _abi.fw_collider:
    """``Collider::cone(radius, height)``, apex towards +Y (examples/textures.rs:211)."""
    return _revolved(_abi.FW_COLLIDER_CONE, radius, height, translation, rotation, layers)


def capsule(radius, length, translation, rotation=(0.0, 0.0, 0.0, 1.0), layers: int = 1) -> _abi.fw_collider:
    """``Collider::capsule(radius, length)``, axis +Y: a segment of ``length`` swept by a ball."""
    return _revolved(_abi.FW_COLLIDER_CAPSULE, radius, length, translation, rotation, layers)


def grid_positions(n: int, spacing: float = 2.0, y: float = 0.1) -> List[Tuple[float, float, float]]:
    """n spawners on a near-square grid (C2: 8x8, C3: 32x16)."""
    cols = int(math.ceil(math.sqrt(n)))
    while n % cols and cols < n:
        cols += 1
    rows = n // cols if n % cols == 0 else int(math.ceil(n / cols))
    out = []
    for i in range(n):
        r, c = divmod(i, cols)
        out.append(((c - (cols - 1) / 2.0) * spacing, y, (r - (rows - 1) / 2.0) * spacing))
    return out


def collision_scene_colliders(n_colliders: int = 256, seed: int = 7) -> List[_abi.fw_collider]:
    """C5: the ground slab of examples/stress_test_collision.rs:85-88 widened to the ring of
    spawners, plus unit cubes rotated rot_x(pi/4)*rot_y(pi/4) (:145-150) at seeded positions."""
    import random

    rnd = random.Random(seed)
    out = [cuboid((24.0, 1.0, 24.0), (0.0, -0.5, 0.0))]
    rot = quat_mul(quat_from_rotation_x(math.pi / 4), quat_from_rotation_y(math.pi / 4))
    while len(out) < n_colliders:
        out.append(cuboid((1.0, 1.0, 1.0), (rnd.uniform(-9.0, 9.0), rnd.uniform(0.5, 6.0), rnd.uniform(-9.0, 9.0)), rot))
    return out


def collision_ring(n_spawners: int = 8, radius: float = 5.0):
    """C5: spawners at radius 5, height 0.5, tilted pi/4 towards the centre (:134-138)."""
    out = []
    for i in range(n_spawners):
        a = 2.0 * math.pi * i / n_spawners
        # the example's spawner sits at (5, .5, 0) rotated about z by pi/4; rotate that about y
        rot = quat_mul(quat_from_rotation_y(-a), quat_from_rotation_z(math.pi / 4))
        out.append(((radius * math.cos(a), 0.5, radius * math.sin(a)), rot))
    return out
