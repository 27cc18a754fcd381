"""Host mirror of the public data model of ``src/core.rs`` (:11-341): the same type and field
names, as plain settings containers plus their conversion to the POD structs of the C ABI.

No simulation arithmetic lives here. ``ParticleSpawnerData.particles`` is a lazily refreshed
host mirror of device state (filled by ``fw_read_particles`` on access), as SURVEY section 8b
prescribes for code like ``examples/stress_test.rs:197-199``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from enum import IntEnum
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import _abi
from .curve import FireworkCurve, FireworkGradient, LinearRgba
from .emission_shape import EmissionShape

Vec3 = Tuple[float, float, float]
Quat = Tuple[float, float, float, float]  # x, y, z, w
QUAT_IDENTITY: Quat = (0.0, 0.0, 0.0, 1.0)


# ---------------------------------------------------------------- bevy_utilitarian 0.10.0
@dataclass(frozen=True)
class RandF32:
    """``bevy_utilitarian::prelude::RandF32`` (used at src/core.rs:102,107,157)."""

    min: float
    max: float

    @staticmethod
    def constant(value: float) -> "RandF32":
        return RandF32(float(value), float(value))

    def to_pod(self) -> _abi.fw_rand_f32:
        return _abi.fw_rand_f32(self.min, self.max)


@dataclass(frozen=True)
class RandVec3:
    """``bevy_utilitarian::prelude::RandVec3`` (used at src/core.rs:155,161)."""

    magnitude: RandF32
    direction: Vec3
    spread: float = 0.0

    @staticmethod
    def constant(value: Vec3) -> "RandVec3":
        x, y, z = (float(c) for c in value)
        length = math.sqrt(x * x + y * y + z * z)
        direction = (x / length, y / length, z / length) if length > 0.0 else (0.0, 0.0, 0.0)
        return RandVec3(RandF32.constant(length), direction, 0.0)

    def to_pod(self) -> _abi.fw_rand_vec3:
        pod = _abi.fw_rand_vec3()
        pod.magnitude = self.magnitude.to_pod()
        pod.direction[:] = [float(c) for c in self.direction]
        pod.spread = float(self.spread)
        return pod


# ---------------------------------------------------------------- src/core.rs:11-97
@dataclass(frozen=True)
class EmissionPacing:
    """``EmissionPacing`` (src/core.rs:11-44)."""

    kind: int
    one_shot_count: int = 0
    count: float = 0.0
    duration: float = 1.0
    offset_start: float = 0.0
    offset_end: float = 1.0

    @staticmethod
    def OneShot(count: int) -> "EmissionPacing":
        return EmissionPacing(_abi.FW_PACING_ONE_SHOT, one_shot_count=int(count))

    @staticmethod
    def CountOverDuration(count: float, duration: float, offset_start: float = 0.0,
                          offset_end: float = 1.0) -> "EmissionPacing":
        return EmissionPacing(_abi.FW_PACING_COUNT_OVER_DURATION, 0, float(count), float(duration),
                              float(offset_start), float(offset_end))

    @staticmethod
    def rate(rate: float) -> "EmissionPacing":
        """src/core.rs:36-43"""
        return EmissionPacing.CountOverDuration(rate, 1.0, 0.0, 1.0)

    def is_one_shot(self) -> bool:
        return self.kind == _abi.FW_PACING_ONE_SHOT


EmissionPacing.OnDemand = EmissionPacing(_abi.FW_PACING_ON_DEMAND)


@dataclass(frozen=True)
class EmissionMode:
    """``EmissionMode`` (src/core.rs:46-54)."""

    kind: int = _abi.FW_MODE_GLOBAL
    target_particle_type: int = 0

    @staticmethod
    def Nested(target_particle_type: int) -> "EmissionMode":
        return EmissionMode(_abi.FW_MODE_NESTED, int(target_particle_type))


EmissionMode.Global = EmissionMode()


class BlendMode(IntEnum):
    """``BlendMode`` (src/core.rs:57-97); render-only, carried for API completeness."""

    Opaque = 0
    Blend = 2
    Premultiplied = 3
    Add = 4
    Multiply = 5


class SpawnTransformMode(IntEnum):
    """``SpawnTransformMode`` (src/core.rs:66-73)."""

    Global = _abi.FW_TRANSFORM_GLOBAL
    Local = _abi.FW_TRANSFORM_LOCAL


@dataclass
class ParticleCollisionSettings:
    """``ParticleCollisionSettings`` (src/core.rs:240-248). ``filter`` is the layer mask of the
    avian ``SpatialQueryFilter`` (default: everything)."""

    restitution: float
    friction: float
    destroy_on_collision: bool = False
    filter: int = 0xFFFFFFFF
    # SpatialQueryFilter::excluded_entities: keys of colliders (fw_collider.key) the sweep does not see
    excluded: Sequence[int] = ()


@dataclass
class ParticleEventHandlers:
    """``ParticleEventHandlers`` (src/core.rs:164-167): callback receives the destroyed rows."""

    particles_destroyed: Optional[Callable[[np.ndarray], None]] = None


@dataclass
class ParticleSettings:
    """``ParticleSettings`` (src/core.rs:99-142) with the defaults of :187-211."""

    lifetime: RandF32 = field(default_factory=lambda: RandF32.constant(5.0))
    scale_curve: FireworkCurve = field(default_factory=lambda: FireworkCurve.constant(1.0))
    initial_scale: RandF32 = field(default_factory=lambda: RandF32.constant(1.0))
    acceleration: Vec3 = (0.0, -9.81, 0.0)
    angular_acceleration: Vec3 = (0.0, 0.0, 0.0)
    linear_drag: float = 0.2
    angular_drag: float = 0.2
    base_color: FireworkGradient = field(default_factory=lambda: FireworkGradient.constant(LinearRgba.WHITE))
    base_color_texture: Optional[object] = None
    emissive_color: FireworkGradient = field(default_factory=lambda: FireworkGradient.constant(LinearRgba.BLACK))
    normal_map_texture: Optional[object] = None
    orm_texture: Optional[object] = None
    fade_edge: float = 0.7
    fade_scene: float = 1.0
    blend_mode: BlendMode = BlendMode.Blend
    pbr: bool = False
    collision_settings: Optional[ParticleCollisionSettings] = None
    event_handlers: ParticleEventHandlers = field(default_factory=ParticleEventHandlers)
    # not in the reference: initial device capacity of this particle type (0 = automatic)
    capacity_hint: int = 0

    def to_pod(self) -> _abi.fw_particle_settings:
        pod = _abi.fw_particle_settings()
        pod.lifetime = self.lifetime.to_pod()
        pod.scale_curve = self.scale_curve.to_pod()
        pod.initial_scale = self.initial_scale.to_pod()
        pod.acceleration[:] = [float(c) for c in self.acceleration]
        pod.angular_acceleration[:] = [float(c) for c in self.angular_acceleration]
        pod.linear_drag = float(self.linear_drag)
        pod.angular_drag = float(self.angular_drag)
        pod.base_color = self.base_color.to_pod()
        pod.emissive_color = self.emissive_color.to_pod()
        pod.pbr = 1 if self.pbr else 0
        cs = self.collision_settings
        if cs is not None:
            pod.collision.enabled = 1
            pod.collision.restitution = float(cs.restitution)
            pod.collision.friction = float(cs.friction)
            pod.collision.destroy_on_collision = 1 if cs.destroy_on_collision else 0
            pod.collision.filter_mask = int(cs.filter) & 0xFFFFFFFF
            if len(cs.excluded) > _abi.FW_MAX_EXCLUDED:
                raise ValueError(f"at most {_abi.FW_MAX_EXCLUDED} excluded colliders per SpatialQueryFilter")
            pod.collision.n_excluded = len(cs.excluded)
            for k, key in enumerate(cs.excluded):
                pod.collision.excluded_keys[k] = int(key) & 0xFFFFFFFF
        pod.capture_destroyed = 1 if self.event_handlers.particles_destroyed is not None else 0
        pod.capacity_hint = int(self.capacity_hint)
        return pod


@dataclass
class EmissionSettings:
    """``EmissionSettings`` (src/core.rs:144-162) with the defaults of :213-227."""

    particle_index: int = 0
    emission_pacing: EmissionPacing = field(default_factory=lambda: EmissionPacing.rate(5.0))
    emission_mode: EmissionMode = EmissionMode.Global
    emission_shape: EmissionShape = EmissionShape.Point
    initial_velocity: RandVec3 = field(default_factory=lambda: RandVec3.constant((0.0, 0.0, 0.0)))
    initial_velocity_radial: RandF32 = field(default_factory=lambda: RandF32.constant(0.0))
    inherit_parent_velocity: bool = True
    initial_rotation: Quat = QUAT_IDENTITY
    initial_angular_velocity: RandVec3 = field(default_factory=lambda: RandVec3.constant((0.0, 0.0, 0.0)))

    def to_pod(self) -> _abi.fw_emission_settings:
        pod = _abi.fw_emission_settings()
        pod.particle_index = int(self.particle_index)
        p = self.emission_pacing
        pod.pacing_kind = p.kind
        pod.one_shot_count = p.one_shot_count
        pod.count, pod.duration = p.count, p.duration
        pod.offset_start, pod.offset_end = p.offset_start, p.offset_end
        pod.mode = self.emission_mode.kind
        pod.target_particle_type = self.emission_mode.target_particle_type
        pod.shape_kind = self.emission_shape.kind
        pod.shape_radius = self.emission_shape.radius
        pod.shape_normal[:] = list(self.emission_shape.normal)
        pod.initial_velocity = self.initial_velocity.to_pod()
        pod.initial_velocity_radial = self.initial_velocity_radial.to_pod()
        pod.inherit_parent_velocity = 1 if self.inherit_parent_velocity else 0
        pod.initial_rotation[:] = [float(c) for c in self.initial_rotation]
        pod.initial_angular_velocity = self.initial_angular_velocity.to_pod()
        return pod


@dataclass
class ParticleSpawner:
    """``ParticleSpawner`` component (src/core.rs:169-185), defaults :229-238."""

    particle_settings: List[ParticleSettings] = field(default_factory=lambda: [ParticleSettings()])
    emission_settings: List[EmissionSettings] = field(default_factory=lambda: [EmissionSettings()])
    starts_enabled: bool = True
    spawn_transform_mode: SpawnTransformMode = SpawnTransformMode.Global

    def pods(self):
        n_t, n_e = len(self.particle_settings), len(self.emission_settings)
        ps = (_abi.fw_particle_settings * max(n_t, 1))()
        es = (_abi.fw_emission_settings * max(n_e, 1))()
        for i, s in enumerate(self.particle_settings):
            ps[i] = s.to_pod()
        for i, s in enumerate(self.emission_settings):
            if not 0 <= s.particle_index < n_t:
                raise IndexError("EmissionSettings.particle_index out of range")
            es[i] = s.to_pod()
        return ps, n_t, es, n_e


@dataclass
class EffectModifier:
    """``EffectModifier`` component (src/core.rs:323-336)."""

    scale: float = 1.0
    speed: float = 1.0


@dataclass
class ParticleSpawnerFinished:
    """``ParticleSpawnerFinished`` entity event (src/core.rs:338-341)."""

    entity: int


# ``ParticleData`` (src/core.rs:305-321) rows come back as a numpy structured array with this
# dtype: fields position, velocity, rotation, angular_velocity, initial_scale, scale, age,
# lifetime, base_color, emissive_color, pbr.
ParticleData = _abi.particle_data_dtype()


class _ParticleVecs(Sequence):
    """``Vec<Vec<ParticleData>>`` view: indexing by particle type reads that stream back."""

    def __init__(self, data: "ParticleSpawnerData"):
        self._data = data

    def __len__(self) -> int:
        return self._data._n_types

    def __getitem__(self, i: int) -> np.ndarray:
        if not 0 <= i < self._data._n_types:
            raise IndexError(i)
        return self._data._engine.read_particles(self._data._key, i)


class ParticleSpawnerData:
    """``ParticleSpawnerData`` component (src/core.rs:269-303) as a handle onto device state."""

    def __init__(self, engine, key: int, n_types: int):
        self._engine = engine
        self._key = key
        self._n_types = n_types
        self.initialized = True
        self.parent_velocity: Vec3 = (0.0, 0.0, 0.0)
        self.manual_queued_count = 0

    @property
    def particles(self) -> _ParticleVecs:
        return _ParticleVecs(self)

    def counts(self) -> List[int]:
        """``data.particles[i].len()`` for every particle type, without copying the rows."""
        return self._engine.counts(self._key, self._n_types)

    def queue_particles(self, count: int) -> None:
        """src/core.rs:284-286"""
        self.manual_queued_count += int(count)

    def active(self) -> bool:
        """src/core.rs:288-302"""
        return bool(self._engine.status(self._key).active)

    @property
    def finished_notified(self) -> bool:
        return bool(self._engine.status(self._key).finished_notified)
