"""SURVEY section 8f ranks 2 and 3 on the GPU: nested emission (particles that emit particles,
reference src/core.rs:471-546) and the destroyed-particle stream handed to
`particles_destroyed` handlers (:164-167, 588, 597, 637, 660-667), against the oracle."""
import math

import numpy as np
import pytest

from bevy_firework_b200 import (EmissionMode, EmissionPacing, EmissionSettings, EmissionShape, FireworkCurve,
                                FireworkGradient, LinearRgba, ParticleCollisionSettings, ParticleEventHandlers,
                                ParticleSettings, ParticleSpawner, RandF32, RandVec3, SpawnTransformMode, _abi)
from bevy_firework_b200._native import frame_input
from bevy_firework_b200.workloads import cuboid
from _parity import assert_rows_match, random_rows, reset_both

pytestmark = pytest.mark.gpu
DT = float(np.float32(1.0) / np.float32(60.0))


def textures_spawner(rate=12.0, nested_count=6.0, parent_lifetime=3.0):
    """examples/textures.rs:67-163: shell casings (type 0) that trail smoke puffs (type 1)."""
    return ParticleSpawner(
        particle_settings=[
            ParticleSettings(lifetime=RandF32.constant(parent_lifetime), initial_scale=RandF32(0.08, 0.1),
                             acceleration=(0.0, -9.81, 0.0), linear_drag=0.1, angular_drag=0.1, pbr=True),
            ParticleSettings(lifetime=RandF32.constant(2.0), scale_curve=FireworkCurve.even_samples([1.0, 2.0]),
                             initial_scale=RandF32(0.5, 0.8), acceleration=(0.0, 0.3, 0.0), linear_drag=0.7,
                             base_color=FireworkGradient.uneven_samples([(0.0, LinearRgba(0.1, 0.1, 0.1, 0.0)),
                                                                         (0.1, LinearRgba(0.1, 0.1, 0.1, 0.15)),
                                                                         (1.0, LinearRgba(0.1, 0.1, 0.1, 0.0))]),
                             emissive_color=FireworkGradient.constant(LinearRgba.BLACK), pbr=True),
        ],
        emission_settings=[
            EmissionSettings(particle_index=0, emission_pacing=EmissionPacing.rate(rate),
                             initial_velocity=RandVec3(RandF32(2.0, 5.0), (0.0, 1.0, 0.0), 0.4),
                             initial_rotation=(0.0, math.sin(math.pi / 4), 0.0, math.cos(math.pi / 4)),
                             initial_angular_velocity=RandVec3(RandF32(5.0, 15.0), (0.0, -1.0, 0.0), 0.0)),
            EmissionSettings(particle_index=1, emission_mode=EmissionMode.Nested(0),
                             emission_pacing=EmissionPacing.CountOverDuration(nested_count, 0.0, 0.0, 0.1),
                             inherit_parent_velocity=False),
        ],
        spawn_transform_mode=SpawnTransformMode.Local)


@pytest.mark.parametrize("rate,nested_count", [(12.0, 6.0), (3000.0, 40.0)])
def test_nested_emission_textures_example(engine, oracle, rate, nested_count):
    sp = textures_spawner(rate, nested_count)
    w = oracle.OracleWorld()
    reset_both(engine, w, 7, sp)
    q = (0.0, 0.0, -math.sin(math.pi / 4), math.cos(math.pi / 4))  # from_rotation_arc(Y, X)
    inp = [frame_input(7, (-2.0, 2.0, 0.0), q)]
    for k in range(240):
        engine.frame(DT, inp)
        w.frame(DT, inp)
        assert engine.counts(7) == w.counts(7), f"frame {k}"
        if k % 80 == 79:
            for t in (0, 1):
                assert_rows_match(engine.read_particles(7, t), w.read_particles(7, t), what=f"frame {k} type {t}")
    engine.sync()
    assert engine.counts(7)[1] > 0
    st, ost = engine.status(7), w.status(7)
    assert (st.active, st.all_empty, st.live_particles) == (ost.active, ost.all_empty, ost.live_particles)


def test_nested_interleaved_with_global_emitters_and_self_target(engine, oracle):
    """emitter order decides the append order: Global(0) -> Nested(1, parents type 0, children
    type 1) -> Global(2, also type 1) -> Nested(3: type 1 particles emit type 1 particles)."""
    sp = ParticleSpawner(
        particle_settings=[ParticleSettings(lifetime=RandF32(0.4, 0.9), linear_drag=0.1),
                           ParticleSettings(lifetime=RandF32.constant(0.5), initial_scale=RandF32(0.1, 0.2))],
        emission_settings=[
            EmissionSettings(particle_index=0, emission_pacing=EmissionPacing.rate(900.0),
                             initial_velocity=RandVec3(RandF32(1.0, 3.0), (0.0, 1.0, 0.0), 0.6)),
            EmissionSettings(particle_index=1, emission_mode=EmissionMode.Nested(0),
                             emission_pacing=EmissionPacing.CountOverDuration(9.0, 0.0, 0.2, 0.9),
                             emission_shape=EmissionShape.Sphere(0.2), inherit_parent_velocity=True),
            EmissionSettings(particle_index=1, emission_pacing=EmissionPacing.rate(500.0),
                             initial_velocity=RandVec3.constant((1.0, 0.0, 0.0))),
            EmissionSettings(particle_index=1, emission_mode=EmissionMode.Nested(1),
                             emission_pacing=EmissionPacing.CountOverDuration(2.0, 0.0, 0.5, 0.6),
                             initial_velocity=RandVec3.constant((0.0, -1.0, 0.0))),
        ])
    w = oracle.OracleWorld()
    reset_both(engine, w, 3, sp)
    inp = [frame_input(3, (0.0, 1.0, 0.0), modifier_scale=1.5, modifier_speed=0.75)]
    for k in range(150):
        engine.frame(DT, inp)
        w.frame(DT, inp)
        assert engine.counts(3) == w.counts(3), f"frame {k}"
    for t in (0, 1):
        assert_rows_match(engine.read_particles(3, t), w.read_particles(3, t), what=f"type {t}")


def test_nested_on_injected_parents(engine, oracle):
    """parents written by the host (fw_write_particles) that are already old: the first nested
    pass emits their whole backlog at once (last_emitted_age starts at f32::MIN, src/core.rs:467)."""
    sp = textures_spawner(rate=0.0, nested_count=7.0)
    sp.emission_settings[0].emission_pacing = EmissionPacing.OneShot(0)
    w = oracle.OracleWorld()
    reset_both(engine, w, 1, sp)
    rows = random_rows(np.random.default_rng(2), 4000, lifetime=(1.0, 3.0))
    rows["pbr"] = 1  # ParticleData.pbr is a copy of the type's setting (src/core.rs:462)
    engine.write_particles(1, 0, rows)
    w.write_particles(1, 0, rows)
    for k in range(20):
        engine.frame(DT, [frame_input(1)])
        w.frame(DT, [frame_input(1)])
        assert engine.counts(1) == w.counts(1), f"frame {k}"
    engine.sync()
    assert engine.counts(1)[1] > 4000
    assert_rows_match(engine.read_particles(1, 1), w.read_particles(1, 1))
    assert_rows_match(engine.read_particles(1, 0), w.read_particles(1, 0))


def test_destroyed_stream_lifetime_deaths(engine, oracle):
    """with a particles_destroyed handler the particles removed by a frame are readable in Vec
    order, age already bumped, nothing else changed (src/core.rs:594-598)."""
    sp = ParticleSpawner(
        particle_settings=[ParticleSettings(lifetime=RandF32(0.3, 2.0), linear_drag=0.2,
                                            base_color=FireworkGradient.even_samples([LinearRgba(1, 0, 0, 1), LinearRgba(0, 0, 1, 0)]),
                                            event_handlers=ParticleEventHandlers(particles_destroyed=lambda rows: None))],
        emission_settings=[EmissionSettings(emission_pacing=EmissionPacing.OneShot(0))])
    w = oracle.OracleWorld()
    reset_both(engine, w, 1, sp)
    rows = random_rows(np.random.default_rng(8), 6000, lifetime=(0.3, 2.0))
    engine.write_particles(1, 0, rows)
    w.write_particles(1, 0, rows)
    total = 0
    for k in range(100):
        engine.frame(DT, [])
        w.frame(DT, [])
        got, want = engine.read_destroyed(1, 0), w.read_destroyed(1, 0)
        assert_rows_match(got, want, what=f"destroyed at frame {k}")
        total += len(got)
    assert total + engine.counts(1)[0] == 6000 and total > 1000
    assert_rows_match(engine.read_particles(1, 0), w.read_particles(1, 0))


def test_destroyed_stream_collision_deaths(engine, oracle):
    """destroy_on_collision: the destroyed record carries the post-collision position, velocity
    and scale but the old colours (src/core.rs:633-639)."""
    sp = ParticleSpawner(
        particle_settings=[ParticleSettings(lifetime=RandF32.constant(50.0),
                                            scale_curve=FireworkCurve.even_samples([1.0, 3.0]),
                                            collision_settings=ParticleCollisionSettings(0.5, 0.1, True),
                                            event_handlers=ParticleEventHandlers(particles_destroyed=lambda rows: None))],
        emission_settings=[EmissionSettings(emission_pacing=EmissionPacing.OneShot(0))])
    cols = [cuboid((8, 1, 8), (0, -0.5, 0))]
    w = oracle.OracleWorld()
    reset_both(engine, w, 1, sp)
    engine.set_colliders(cols)
    w.set_colliders(cols)
    rng = np.random.default_rng(4)
    rows = random_rows(rng, 3000, angular=False)
    rows["position"][:, 1] = rng.uniform(0.05, 2.0, 3000)
    rows["lifetime"] = 50.0
    rows["age"] = rng.uniform(0.0, 20.0, 3000).astype(np.float32)
    engine.write_particles(1, 0, rows)
    w.write_particles(1, 0, rows)
    total = 0
    for k in range(40):
        engine.frame(DT, [])
        w.frame(DT, [])
        got, want = engine.read_destroyed(1, 0), w.read_destroyed(1, 0)
        assert_rows_match(got, want, what=f"frame {k}")
        total += len(got)
    assert 100 < total < 3000


def test_destroyed_handler_through_the_plugin():
    """ParticleEventHandlers.particles_destroyed is called with the destroyed rows each frame."""
    from bevy_firework_b200.plugin import App, ParticleSystemPlugin, Transform

    seen = []
    sp = ParticleSpawner(
        particle_settings=[ParticleSettings(lifetime=RandF32(0.1, 0.3),
                                            event_handlers=ParticleEventHandlers(particles_destroyed=lambda rows: seen.append(len(rows))))],
        emission_settings=[EmissionSettings(emission_pacing=EmissionPacing.OneShot(200))])
    app = App().add_plugins(ParticleSystemPlugin(device=0))
    e = app.spawn(sp, Transform())
    for _ in range(25):
        app.update(DT)
    assert sum(seen) == 200 and app.data(e).counts() == [0]


def test_textures_example_with_its_colliders(engine, oracle):
    """examples/textures.rs as a whole: the casings (type 0) bounce off the circular base
    (Collider::cylinder(4, 0.2), :195) and the cone (:211) with the example's collision settings
    (:97-102) while they trail nested smoke puffs (type 1)."""
    from bevy_firework_b200.workloads import cone, cylinder

    sp = textures_spawner(rate=600.0, nested_count=10.0)
    sp.particle_settings[0].collision_settings = ParticleCollisionSettings(0.4, 0.35, False)
    cols = [cylinder(4.0, 0.2, (0.0, 0.0, 0.0)), cone(0.5, 1.0, (0.0, 0.5, 0.0))]
    w = oracle.OracleWorld()
    engine.set_colliders(cols)
    w.set_colliders(cols)
    reset_both(engine, w, 7, sp)
    q = (0.0, 0.0, -math.sin(math.pi / 4), math.cos(math.pi / 4))  # from_rotation_arc(Y, X)
    inp = [frame_input(7, (-2.0, 2.0, 0.0), q)]
    for k in range(200):
        engine.frame(DT, inp)
        w.frame(DT, inp)
        assert engine.counts(7) == w.counts(7), f"frame {k}"
    got, want = engine.read_particles(7, 0), w.read_particles(7, 0)
    assert len(got) == len(want) > 1000
    assert_rows_match(got, want, what="casings")
    # most casings end up resting on the base (top at y = 0.1), none fell through it inside its radius
    on_disc = np.hypot(want["position"][:, 0], want["position"][:, 2]) < 3.9
    assert (want["position"][on_disc, 1] > 0.05).all()
    puffs, opuffs = engine.read_particles(7, 1), w.read_particles(7, 1)
    assert len(puffs) == len(opuffs) > 0
    assert (puffs["age"] == opuffs["age"]).all()
