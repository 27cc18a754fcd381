"""The plugin-level API (host mirror of src/plugin.rs / src/core.rs) on the GPU: scenes written
like the reference's examples, checked against the oracle playing the same schedule."""
import math

import numpy as np
import pytest

from bevy_firework_b200 import (EffectModifier, EmissionPacing, EmissionSettings, ParticleSettings,
                                ParticleSpawner, RandF32, RandVec3, SpawnTransformMode)
from bevy_firework_b200._native import frame_input
from bevy_firework_b200.plugin import App, ParticleSystemPlugin, Transform
from bevy_firework_b200.workloads import one_shot_spawner, sparks_spawner
from _parity import assert_rows_match

pytestmark = pytest.mark.gpu
DT = float(np.float32(1.0) / np.float32(60.0))


def test_sparks_example_through_the_plugin(oracle):
    """examples/sparks.rs: App + ParticleSystemPlugin + one spawner at (0, 0.1, 0)."""
    app = App().add_plugins(ParticleSystemPlugin(device=0, seed=0x00F12E00))
    e = app.spawn(sparks_spawner(), Transform.from_xyz(0.0, 0.1, 0.0))
    w = oracle.OracleWorld(seed=0x00F12E00)
    ps, nt, es, ne = sparks_spawner().pods()
    w.spawner_reset(e, ps, nt, es, ne, True)
    for _ in range(120):
        app.update(DT)
        w.frame(DT, [frame_input(e, (0.0, 0.1, 0.0))])
    data = app.data(e)
    assert data.counts() == w.counts(e)
    assert len(data.particles) == 1
    rows = data.particles[0]                      # lazily mirrored Vec<ParticleData>
    assert_rows_match(rows, w.read_particles(e, 0))
    assert data.active()


def test_one_shot_finished_observer_and_despawn():
    """examples/one_shot.rs:137-141: observe ParticleSpawnerFinished, despawn the entity."""
    app = App().add_plugins(ParticleSystemPlugin(device=0))
    finished = []
    e = app.spawn(one_shot_spawner(count=500, lifetime=0.25), Transform.from_xyz(1.0, 2.0, 3.0))
    app.observe(e, lambda ev: (finished.append(ev.entity), app.despawn(ev.entity)))
    for k in range(1, 30):
        app.update(DT)
        if finished:
            break
    assert finished == [e] and k == 15            # lifetime 0.25 s: the f32 age sum reaches it on update #15
    app.update(DT)                                # the despawned entity is gone from the engine too
    assert app.engine.total_live() == 0


def test_modifier_propagates_to_descendants_and_local_transform(oracle):
    """propagate_particle_spawner_modifier (src/core.rs:690-703) + SpawnTransformMode (:432-435)."""
    sp = ParticleSpawner(
        particle_settings=[ParticleSettings(lifetime=RandF32.constant(2.0), initial_scale=RandF32(0.5, 1.0))],
        emission_settings=[EmissionSettings(emission_pacing=EmissionPacing.rate(600.0),
                                            initial_velocity=RandVec3.constant((0.0, 2.0, 0.0)))],
        spawn_transform_mode=SpawnTransformMode.Global)
    app = App().add_plugins(ParticleSystemPlugin(device=0, seed=99))
    q = (0.0, 0.0, math.sin(0.25), math.cos(0.25))
    root = app.spawn(None, Transform((5.0, 0.0, 0.0), q), modifier=EffectModifier(scale=2.0, speed=0.5))
    child = app.spawn(sp, Transform.from_xyz(0.0, 1.0, 0.0), parent=root)
    for _ in range(30):
        app.update(DT)
    w = oracle.OracleWorld(seed=99)
    ps, nt, es, ne = sp.pods()
    w.spawner_reset(child, ps, nt, es, ne, True)
    g = Transform((5.0, 0.0, 0.0), q).mul_transform(Transform.from_xyz(0.0, 1.0, 0.0))
    for _ in range(30):
        w.frame(DT, [frame_input(child, g.translation, g.rotation, (0, 0, 0), 2.0, 0.5)])
    rows = app.data(child).particles[0]
    assert_rows_match(rows, w.read_particles(child, 0))
    assert rows["initial_scale"].min() >= 1.0      # scaled by the inherited modifier


def test_changed_spawner_resets_particles():
    """mutating the ParticleSpawner component re-runs sync_spawner_data: particles are dropped."""
    app = App().add_plugins(ParticleSystemPlugin(device=0))
    e = app.spawn(sparks_spawner(3000.0), Transform())
    for _ in range(20):
        app.update(DT)
    assert app.data(e).counts()[0] > 900
    app.spawner_mut(e).particle_settings[0].linear_drag = 0.5
    app.update(DT)
    assert app.data(e).counts()[0] == 50           # only this frame's emission
