"""Host mirror of ``src/plugin.rs``: ``ParticleSystemPlugin`` and the system chain it installs
(:46-60), played by a minimal ``App`` that stands in for the Bevy schedule in this harness
(no Rust toolchain exists here; the Rust shim is in ``rust/`` and INTEGRATION.md).

Per ``App.update(dt)`` the chain is, in the reference's order:

1. ``propagate_particle_spawner_modifier`` (src/core.rs:690-703)  -- host, on the entity tree
2. ``sync_spawner_data`` for changed spawners (:343-365)          -- ``fw_spawner_reset``
3. ``sync_parent_velocity`` (:705-742)                            -- host, user-provided velocity
4. ``spawn_particles`` ; ``update_particles`` (:367-670)          -- ONE ``fw_frame`` call
5. ``notify_finished_particle_spawners`` (:674-688)               -- ``fw_spawner_status_get``

Everything per-particle happens on the GPU behind the C ABI; this file only owns entity
bookkeeping.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

from . import _abi
from ._native import Engine, frame_input
from .core import (EffectModifier, ParticleSpawner, ParticleSpawnerData, ParticleSpawnerFinished,
                   SpawnTransformMode)

Vec3 = Tuple[float, float, float]
Quat = Tuple[float, float, float, float]


def _quat_mul(a: Quat, b: Quat) -> Quat:
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return (aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
            aw * bw - ax * bx - ay * by - az * bz)


def _quat_rotate(q: Quat, v: Vec3) -> Vec3:
    x, y, z, w = q
    b2 = x * x + y * y + z * z
    d = v[0] * x + v[1] * y + v[2] * z
    cx, cy, cz = (y * v[2] - z * v[1], z * v[0] - x * v[2], x * v[1] - y * v[0])
    k, m, n = w * w - b2, 2.0 * d, 2.0 * w
    return (v[0] * k + x * m + cx * n, v[1] * k + y * m + cy * n, v[2] * k + z * m + cz * n)


@dataclass
class Transform:
    """bevy ``Transform`` (translation + rotation; scale is not used by the particle path)."""

    translation: Vec3 = (0.0, 0.0, 0.0)
    rotation: Quat = (0.0, 0.0, 0.0, 1.0)

    @staticmethod
    def from_xyz(x: float, y: float, z: float) -> "Transform":
        return Transform((float(x), float(y), float(z)))

    @staticmethod
    def from_translation(t: Vec3) -> "Transform":
        return Transform(tuple(float(c) for c in t))

    def mul_transform(self, child: "Transform") -> "Transform":
        r = _quat_rotate(self.rotation, child.translation)
        return Transform((self.translation[0] + r[0], self.translation[1] + r[1], self.translation[2] + r[2]),
                         _quat_mul(self.rotation, child.rotation))


@dataclass
class _Entity:
    id: int
    spawner: Optional[ParticleSpawner] = None
    transform: Transform = field(default_factory=Transform)
    parent: Optional[int] = None
    modifier: Optional[EffectModifier] = None
    data: Optional[ParticleSpawnerData] = None
    changed: bool = True
    observers: List[Callable[[ParticleSpawnerFinished], None]] = field(default_factory=list)


class ParticleSystemPlugin:
    """``ParticleSystemPlugin`` (src/plugin.rs:22-32). ``update_schedule`` is kept for API
    parity; the extra arguments pick the GPU and the RNG seed of the spawn protocol."""

    def __init__(self, update_schedule: str = "Update", device: int = 0, seed: int = 0x00F12E00,
                 profile: bool = False):
        self.update_schedule = update_schedule
        self.device = device
        self.seed = seed
        self.profile = profile

    def build(self, app: "App") -> None:
        """src/plugin.rs:35-61"""
        app._engine = Engine(device=self.device, seed=self.seed, profile=self.profile)
        app._plugin = self


class App:
    """Stand-in for ``bevy::app::App``: entities with (ParticleSpawner, Transform, optional
    parent / EffectModifier) and the plugin's system chain."""

    def __init__(self):
        self._engine: Optional[Engine] = None
        self._plugin: Optional[ParticleSystemPlugin] = None
        self._entities: Dict[int, _Entity] = {}
        self._next_id = 1

    # -- app / world API
    def add_plugins(self, plugin: ParticleSystemPlugin) -> "App":
        plugin.build(self)
        return self

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            raise RuntimeError("add_plugins(ParticleSystemPlugin()) first")
        return self._engine

    def spawn(self, spawner: Optional[ParticleSpawner] = None, transform: Optional[Transform] = None,
              parent: Optional[int] = None, modifier: Optional[EffectModifier] = None) -> int:
        eid = self._next_id
        self._next_id += 1
        self._entities[eid] = _Entity(eid, spawner, transform or Transform(), parent, modifier)
        return eid

    def despawn(self, entity: int) -> None:
        e = self._entities.pop(entity, None)
        if e is not None and e.data is not None:
            self.engine.spawner_remove(entity)

    def observe(self, entity: int, callback: Callable[[ParticleSpawnerFinished], None]) -> None:
        """``.observe(|trigger: On<ParticleSpawnerFinished>| ...)`` (examples/one_shot.rs:137-141)"""
        self._entities[entity].observers.append(callback)

    def spawner_mut(self, entity: int) -> ParticleSpawner:
        """``Mut<ParticleSpawner>``: marks the component changed (src/core.rs:344)."""
        e = self._entities[entity]
        e.changed = True
        return e.spawner

    def transform_mut(self, entity: int) -> Transform:
        return self._entities[entity].transform

    def data(self, entity: int) -> ParticleSpawnerData:
        d = self._entities[entity].data
        if d is None:
            raise KeyError("spawner data is created by the first update (sync_spawner_data)")
        return d

    def insert_modifier(self, entity: int, modifier: EffectModifier) -> None:
        self._entities[entity].modifier = modifier

    def set_parent_velocity(self, entity: int, velocity: Vec3) -> None:
        """what ``sync_parent_velocity`` (src/core.rs:705-742) writes"""
        self.data(entity).parent_velocity = tuple(float(c) for c in velocity)

    # -- helpers
    def _global_transform(self, e: _Entity) -> Transform:
        chain = []
        cur: Optional[_Entity] = e
        while cur is not None:
            chain.append(cur.transform)
            cur = self._entities.get(cur.parent) if cur.parent is not None else None
        t = chain.pop()
        while chain:
            t = t.mul_transform(chain.pop())
        return t

    def _descendants(self, root: int):
        stack = [k for k, v in self._entities.items() if v.parent == root]
        while stack:
            k = stack.pop()
            yield k
            stack.extend(c for c, v in self._entities.items() if v.parent == k)

    # -- the schedule
    def update(self, dt: float) -> None:
        eng = self.engine
        # propagate_particle_spawner_modifier (src/core.rs:690-703)
        for eid, e in list(self._entities.items()):
            if e.modifier is not None:
                for child in self._descendants(eid):
                    if self._entities[child].spawner is not None:
                        self._entities[child].modifier = e.modifier
        # sync_spawner_data for Changed<ParticleSpawner> (src/core.rs:343-365)
        for eid, e in self._entities.items():
            if e.spawner is not None and e.changed:
                ps, n_t, es, n_e = e.spawner.pods()
                eng.spawner_reset(eid, ps, n_t, es, n_e, e.spawner.starts_enabled)
                if e.data is None:
                    e.data = ParticleSpawnerData(eng, eid, n_t)
                else:
                    e.data._n_types = n_t
                e.changed = False
        # spawn_particles ; update_particles -> one batched call
        inputs = []
        for eid, e in self._entities.items():
            if e.spawner is None:
                continue
            origin = (self._global_transform(e) if e.spawner.spawn_transform_mode == SpawnTransformMode.Global
                      else e.transform)  # src/core.rs:432-435
            mod = e.modifier or EffectModifier()
            inputs.append(frame_input(eid, origin.translation, origin.rotation, e.data.parent_velocity,
                                      mod.scale, mod.speed, e.data.manual_queued_count))
            e.data.manual_queued_count = 0
        eng.frame(dt, inputs)
        # particles_destroyed handlers (src/core.rs:660-667): called with the non-empty destroyed Vec
        for eid, e in list(self._entities.items()):
            if e.spawner is None:
                continue
            for t, ps in enumerate(e.spawner.particle_settings):
                handler = ps.event_handlers.particles_destroyed
                if handler is not None:
                    rows = eng.read_destroyed(eid, t)
                    if len(rows):
                        handler(rows)
        # notify_finished_particle_spawners (src/core.rs:674-688); only one-shot style spawners
        # can finish, so the status is only read back for entities that have observers
        for eid, e in list(self._entities.items()):
            if e.spawner is None or not e.observers:
                continue
            st = eng.status(eid)
            if st.finished:
                eng.mark_finished_notified(eid)
                for cb in list(e.observers):
                    cb(ParticleSpawnerFinished(eid))
