"""What the library keeps per particle (fw_stream_layout_get) and that it never shows: a stream
whose rotation / angular velocity / emissive colour / scale factor / lifetime are provably constant
moves 80 bytes per particle instead of 156, and every read-back still equals the oracle's rows --
including after host-written rows break a proof and the field becomes per-particle state again."""
import numpy as np
import pytest

from bevy_firework_b200 import (EmissionPacing, EmissionSettings, FireworkCurve, FireworkGradient, LinearRgba,
                                ParticleCollisionSettings, ParticleSettings, ParticleSpawner, RandF32, RandVec3, _abi)
from bevy_firework_b200._native import FireworkError, frame_input
from bevy_firework_b200.workloads import cuboid, one_shot_spawner, sparks_spawner, stress_spawner
from _parity import assert_rows_match, random_rows, reset_both

pytestmark = pytest.mark.gpu
DT = float(np.float32(1.0) / np.float32(60.0))
ROT, COMPACT, COLLIDE = _abi.FW_LAYOUT_ROTATES, _abi.FW_LAYOUT_COMPACTING, _abi.FW_LAYOUT_COLLIDES


def _layout(engine, spawner, key=1):
    ps, nt, es, ne = spawner.pods()
    engine.spawner_reset(key, ps, nt, es, ne, True)
    return engine.stream_layout(key, 0)


def test_layout_of_the_baseline_configs(engine):
    lay = _layout(engine, stress_spawner(rate=1000.0))  # C2 / C3: nothing but position, age, velocity, colour varies
    assert (lay.variant, lay.flags) == (0, _abi.FW_STORE_BASE_COLOR) and (lay.bytes_read, lay.bytes_written) == (32, 48)
    lay = _layout(engine, sparks_spawner(1000.0), 2)
    assert (lay.bytes_read, lay.bytes_written) == (32, 48)
    lay = _layout(engine, one_shot_spawner(1000, 2.5), 3)  # C4: the scale curve is not constant
    assert lay.flags == _abi.FW_STORE_BASE_COLOR | _abi.FW_STORE_SCALE and (lay.bytes_read, lay.bytes_written) == (32, 52)
    lay = _layout(engine, stress_spawner(rate=1000.0, lifetime=1.0, lifetime_spread=0.5), 4)  # C3r
    assert lay.variant == COMPACT and lay.flags & _abi.FW_STORE_LIFETIME
    assert (lay.bytes_read, lay.bytes_written, lay.bytes_count_pass) == (40, 56, 8)


def test_layout_of_a_stream_where_nothing_is_provable(engine):
    sp = ParticleSpawner(
        particle_settings=[ParticleSettings(lifetime=RandF32(0.5, 1.5), scale_curve=FireworkCurve.even_samples([1.0, 0.0]),
                                            base_color=FireworkGradient.even_samples([LinearRgba(1, 0, 0, 1), LinearRgba(0, 0, 1, 0)]),
                                            emissive_color=FireworkGradient.even_samples([LinearRgba(1, 1, 1, 1), LinearRgba(0, 0, 0, 0)]))],
        emission_settings=[EmissionSettings(emission_pacing=EmissionPacing.rate(100.0),
                                            initial_angular_velocity=RandVec3(RandF32(1.0, 2.0), (0.0, 1.0, 0.0), 0.3))])
    lay = _layout(engine, sp)
    assert lay.variant == COMPACT | ROT and lay.flags == 15
    assert (lay.bytes_read, lay.bytes_written, lay.bytes_count_pass) == (64, 100, 24)  # 64 + 92, + the 8 constants compaction moves


@pytest.mark.parametrize("what", ["angular_acceleration", "negative_zero_acceleration", "two_rotations", "handler"])
def test_no_proof_no_shortcut(engine, what):
    """each of these defeats the 'static' proof: the stream keeps rotation and angular velocity"""
    ps = dict(lifetime=RandF32.constant(1.0))
    es = [EmissionSettings(emission_pacing=EmissionPacing.rate(100.0))]
    if what == "angular_acceleration":
        ps["angular_acceleration"] = (0.0, 0.5, 0.0)
    elif what == "negative_zero_acceleration":
        ps["angular_acceleration"] = (0.0, -0.0, 0.0)  # -0 - (+-0 * drag) keeps a data-dependent sign
    elif what == "two_rotations":
        es.append(EmissionSettings(emission_pacing=EmissionPacing.rate(50.0), initial_rotation=(0.0, 0.70710677, 0.0, 0.70710677)))
    sp = ParticleSpawner(particle_settings=[ParticleSettings(**ps)], emission_settings=es)
    if what == "handler":
        from bevy_firework_b200 import ParticleEventHandlers

        sp.particle_settings[0].event_handlers = ParticleEventHandlers(particles_destroyed=lambda rows: None)
    lay = _layout(engine, sp)
    assert lay.variant & ROT


def test_static_stream_reads_back_like_the_oracle(engine, oracle):
    """spawned rows of a static stream: rotation = identity * initial_rotation, angular velocity +0,
    emissive = the constant, scale = initial_scale * constant -- all synthesised, all equal to the
    oracle's stored values; also with a non-identity initial rotation and a constant scale curve != 1"""
    sp = stress_spawner(rate=3000.0)
    sp.particle_settings[0].scale_curve = FireworkCurve.constant(1.7)
    sp.particle_settings[0].emissive_color = FireworkGradient.constant(LinearRgba(0.25, 0.5, 0.75, 1.0))
    sp.emission_settings[0].initial_rotation = (0.0, 0.38268343, 0.0, 0.9238795)
    sp.emission_settings[0].initial_angular_velocity = RandVec3(RandF32(0.0, 0.0), (0.0, -1.0, 0.0), 0.4)  # -0 components at spawn
    w = oracle.OracleWorld()
    reset_both(engine, w, 1, sp)
    assert not engine.stream_layout(1, 0).variant & ROT
    inp = [frame_input(1, (0.0, 0.1, 0.0))]
    for k in range(90):
        engine.frame(DT, inp)
        w.frame(DT, inp)
        if k % 30 == 29:
            assert_rows_match(engine.read_particles(1, 0), w.read_particles(1, 0), what=f"frame {k}")
            gi, rows = engine.read_instances(1, 0), w.read_particles(1, 0)
            for f in ("position", "scale", "rotation", "base_color", "emissive_color"):
                assert (gi[f] == rows[f]).all(), f


def test_host_rows_that_break_a_proof_become_state(engine, oracle):
    """fw_write_particles with rows that contradict the constants: the stream turns the field on and
    keeps matching the oracle; rows that agree with them leave the layout alone"""
    sp = stress_spawner(rate=2000.0)
    w = oracle.OracleWorld()
    reset_both(engine, w, 1, sp)
    inp = [frame_input(1, (0.0, 0.1, 0.0))]
    for _ in range(20):
        engine.frame(DT, inp)
        w.frame(DT, inp)
    rows = engine.read_particles(1, 0)
    engine.write_particles(1, 0, rows)  # its own rows: every proof still holds
    w.write_particles(1, 0, rows)
    lay = engine.stream_layout(1, 0)
    assert (lay.variant, lay.flags) == (0, _abi.FW_STORE_BASE_COLOR)
    assert_rows_match(engine.read_particles(1, 0), rows, what="round trip")
    for _ in range(5):
        engine.frame(DT, inp)
        w.frame(DT, inp)
    assert_rows_match(engine.read_particles(1, 0), w.read_particles(1, 0), what="after round trip")
    # now rows with angular velocity, other emissive colours, odd scales and lifetimes
    rng = np.random.default_rng(11)
    bad = random_rows(rng, 3000, lifetime=(0.5, 2.0))
    bad["scale"] *= 1.5  # (random_rows leaves scale = initial_scale, which the constant curve 1.0 would explain)
    engine.write_particles(1, 0, bad)
    w.write_particles(1, 0, bad)
    lay = engine.stream_layout(1, 0)
    assert lay.variant == ROT | COMPACT and lay.flags == 15
    assert_rows_match(engine.read_particles(1, 0), bad, what="written rows")
    for k in range(40):
        engine.frame(DT, inp)
        w.frame(DT, inp)
        assert engine.counts(1) == w.counts(1), k
    assert_rows_match(engine.read_particles(1, 0), w.read_particles(1, 0), what="after breaking the proofs")


@pytest.mark.parametrize("break_field", ["rotation", "angular_velocity", "emissive_color", "scale", "lifetime", "age_order"])
def test_each_proof_breaks_alone(engine, oracle, break_field):
    sp = stress_spawner(rate=1500.0)
    w = oracle.OracleWorld()
    reset_both(engine, w, 1, sp)
    inp = [frame_input(1, (0.0, 0.1, 0.0))]
    for _ in range(12):
        engine.frame(DT, inp)
        w.frame(DT, inp)
    rows = engine.read_particles(1, 0).copy()
    i = len(rows) // 2
    if break_field == "rotation":
        rows["rotation"][i] = (0.0, 1.0, 0.0, 0.0)
    elif break_field == "angular_velocity":
        rows["angular_velocity"][i] = (0.0, 3.0, 0.0)
    elif break_field == "emissive_color":
        rows["emissive_color"][i] = (1.0, 2.0, 3.0, 4.0)
    elif break_field == "scale":
        rows["scale"][i] *= 3.0
    elif break_field == "lifetime":
        rows["lifetime"][i] = 0.3
    else:
        rows["age"][i] += 0.05
    engine.write_particles(1, 0, rows)
    w.write_particles(1, 0, rows)
    lay = engine.stream_layout(1, 0)
    want_variant = {"rotation": ROT, "angular_velocity": ROT, "lifetime": COMPACT, "age_order": COMPACT}.get(break_field, 0)
    want_flags = _abi.FW_STORE_BASE_COLOR | {"emissive_color": _abi.FW_STORE_EMISSIVE_COLOR, "scale": _abi.FW_STORE_SCALE,
                                             "lifetime": _abi.FW_STORE_LIFETIME}.get(break_field, 0)
    assert (lay.variant, lay.flags) == (want_variant, want_flags)
    assert_rows_match(engine.read_particles(1, 0), rows, what="written rows")
    for k in range(70):
        engine.frame(DT, inp)
        w.frame(DT, inp)
        assert engine.counts(1) == w.counts(1), k
        if k in (0, 30, 69):
            assert_rows_match(engine.read_particles(1, 0), w.read_particles(1, 0), what=f"{break_field} frame {k}")


def test_static_compacting_streams(engine, oracle):
    """static streams in the compacting variants: random lifetimes (k = lifetime + copy of age, the
    8-byte counting pass), and destroy_on_collision with one lifetime (look-back, no lifetime pack)"""
    w = oracle.OracleWorld()
    a = stress_spawner(rate=4000.0, lifetime=0.6, lifetime_spread=0.4)
    b = stress_spawner(rate=4000.0)
    b.particle_settings[0].collision_settings = ParticleCollisionSettings(restitution=0.5, friction=0.2, destroy_on_collision=True)
    cols = [cuboid((30.0, 1.0, 30.0), (0.0, -0.5, 0.0)), cuboid((1.0, 1.0, 1.0), (0.3, 1.5, 0.0))]
    engine.set_colliders(cols)
    w.set_colliders(cols)
    reset_both(engine, w, 1, a)
    reset_both(engine, w, 2, b)
    la, lb = engine.stream_layout(1, 0), engine.stream_layout(2, 0)
    assert la.variant == COMPACT and la.bytes_count_pass == 8
    assert lb.variant == COMPACT | COLLIDE and not lb.flags & _abi.FW_STORE_LIFETIME
    inp = [frame_input(1, (0.0, 0.1, 0.0)), frame_input(2, (0.0, 0.6, 0.0))]
    for k in range(100):
        engine.frame(DT, inp)
        w.frame(DT, inp)
        assert engine.counts(1) == w.counts(1) and engine.counts(2) == w.counts(2), k
    for key in (1, 2):
        assert_rows_match(engine.read_particles(key, 0), w.read_particles(key, 0), what=f"spawner {key}")
    # the ring grows while it runs: a capacity hint far too small
    c = stress_spawner(rate=6000.0, lifetime=0.6, lifetime_spread=0.4)
    c.particle_settings[0].capacity_hint = 1024
    reset_both(engine, w, 3, c)
    inp.append(frame_input(3, (1.0, 0.1, 0.0)))
    for k in range(80):
        engine.frame(DT, inp)
        w.frame(DT, inp)
    assert engine.stream_layout(3, 0).capacity > 4096
    assert_rows_match(engine.read_particles(3, 0), w.read_particles(3, 0), what="grown ring")


def test_edge_hit_on_the_device(engine, oracle):
    """particles aimed exactly at a cuboid's edge (tests/test_oracle_golden.py::test_cuboid_edge_hit_normal):
    finite, and bit-equal to the oracle"""
    sp = ParticleSpawner(particle_settings=[ParticleSettings(lifetime=RandF32.constant(5.0),
                                                             collision_settings=ParticleCollisionSettings(restitution=0.5, friction=0.1))],
                         emission_settings=[EmissionSettings(emission_pacing=EmissionPacing.OneShot(0))])
    cols = [cuboid((1.0, 1.0, 1.0), (0.0, 0.0, 0.0))]
    engine.set_colliders(cols)
    w = oracle.OracleWorld()
    w.set_colliders(cols)
    reset_both(engine, w, 1, sp)
    rows = random_rows(np.random.default_rng(2), 64, lifetime=(5.0, 5.0), angular=False)
    rows["age"] = 0.0
    rows["lifetime"] = 5.0
    rows["position"] = (-0.55, -0.55, 0.0)
    rows["velocity"] = (6.0, 6.0, 0.0)
    rows["position"][32:] = (0.55, 0.55, 0.55)  # corner
    rows["velocity"][32:] = (-6.0, -6.0, -6.0)
    engine.write_particles(1, 0, rows)
    w.write_particles(1, 0, rows)
    engine.frame(DT, [])
    w.frame(DT, [])
    got = engine.read_particles(1, 0)
    assert np.isfinite(got["position"]).all() and np.isfinite(got["velocity"]).all()
    assert (got["velocity"][:32, :2] < 0).all() and (got["velocity"][32:] > 0).all()
    assert_rows_match(got, w.read_particles(1, 0), what="edge / corner hits")


def test_dt_must_be_a_duration(engine):
    """Res<Time>::delta_secs() is finite and >= 0 (src/core.rs:413,594); anything else is refused"""
    ps, nt, es, ne = stress_spawner(rate=100.0).pods()
    engine.spawner_reset(1, ps, nt, es, ne, True)
    for bad in (float("nan"), float("inf"), -DT, -0.0):
        with pytest.raises(FireworkError) as e:
            engine.frame(bad, [])
        assert e.value.code == _abi.FW_ERR_INVALID_ARGUMENT
    engine.frame(0.0, [])  # a paused clock is fine
    engine.frame(DT, [])
    assert engine.counts(1) == [1]


def test_poll_device_errors_never_waits_and_reports_nothing_on_a_healthy_run(engine):
    ps, nt, es, ne = stress_spawner(rate=5000.0).pods()
    engine.spawner_reset(1, ps, nt, es, ne, True)
    flags = 0
    for _ in range(30):
        engine.frame(DT, [frame_input(1, (0.0, 0.1, 0.0))])
        flags |= engine.poll_device_errors()
    engine.sync()
    assert flags == 0 and engine.poll_device_errors() == 0
