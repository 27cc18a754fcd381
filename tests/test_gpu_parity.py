"""Parity of the CUDA path (through the C ABI) against the CPU oracle. GPU only.

Bar: counts exact every frame and EVERY field of every row equal, no tolerance anywhere
(tests/_parity.py) -- spawned or injected state, with or without collisions."""
import math

import numpy as np
import pytest

from bevy_firework_b200 import (EmissionPacing, EmissionSettings, EmissionShape, FireworkCurve,
                                FireworkGradient, LinearRgba, ParticleCollisionSettings,
                                ParticleSettings, ParticleSpawner, RandF32, RandVec3, _abi)
from bevy_firework_b200._native import frame_input
from bevy_firework_b200.workloads import (collision_ring, collision_scene_colliders, collision_spawner,
                                          cuboid, grid_positions, one_shot_spawner, sparks_spawner,
                                          sphere, stress_spawner)
from _parity import assert_rows_match, random_rows, reset_both

pytestmark = pytest.mark.gpu
f32 = np.float32
DT = float(f32(1.0) / f32(60.0))


def _idle_spawner(**ps_kwargs):
    """a spawner that never emits: state is injected with write_particles"""
    return ParticleSpawner(particle_settings=[ParticleSettings(**ps_kwargs)],
                           emission_settings=[EmissionSettings(emission_pacing=EmissionPacing.OneShot(0))])


CURVES = {
    "constant": (FireworkCurve.constant(1.0), FireworkGradient.constant(LinearRgba(0.3, 0.4, 0.5, 1.0))),
    "even": (FireworkCurve.even_samples([1.0, 2.0, 0.5]),
             FireworkGradient.even_samples([LinearRgba(1, 0, 0, 1), LinearRgba(0, 1, 0, 1), LinearRgba(0, 0, 1, 0)])),
    "uneven": (FireworkCurve.uneven_samples([(0.0, 0.2), (0.3, 1.0), (1.0, 0.0)]),
               stress_spawner().particle_settings[0].base_color),
}


@pytest.mark.parametrize("kind", ["constant", "even", "uneven"])
@pytest.mark.parametrize("n", [1, 31, 257, 10037])
def test_single_step_prefilled_bitexact(engine, oracle, kind, n):
    """R2 on identical input state: everything but the rotation quaternion is bit-exact."""
    curve, grad = CURVES[kind]
    sp = _idle_spawner(lifetime=RandF32(0.5, 3.0), scale_curve=curve, base_color=grad, emissive_color=grad,
                       linear_drag=0.1, angular_drag=0.3, angular_acceleration=(0.1, -0.2, 0.3))
    w = oracle.OracleWorld()
    reset_both(engine, w, 5, sp)
    rows = random_rows(np.random.default_rng(n), n)
    engine.write_particles(5, 0, rows)
    w.write_particles(5, 0, rows)
    assert_rows_match(engine.read_particles(5, 0), rows, what="write/read roundtrip")
    engine.frame(DT, [])
    w.frame(DT, [])
    got, want = engine.read_particles(5, 0), w.read_particles(5, 0)
    assert_rows_match(got, want, what=f"{kind} n={n}")
    inst = engine.read_instances(5, 0)
    for f in ("position", "scale", "rotation", "base_color", "emissive_color"):
        assert (inst[f] == got[f]).all()          # ParticleInstance row == From<&ParticleData> (src/render.rs:105-115)


def test_600_steps_trajectory_with_deaths(engine, oracle):
    """random lifetimes -> compact variant: survivors keep the reference's Vec order."""
    sp = _idle_spawner(lifetime=RandF32(0.5, 8.0), base_color=CURVES["uneven"][1], linear_drag=0.2)
    w = oracle.OracleWorld()
    reset_both(engine, w, 1, sp)
    rows = random_rows(np.random.default_rng(3), 5000, lifetime=(0.5, 8.0))
    engine.write_particles(1, 0, rows)
    w.write_particles(1, 0, rows)
    for k in range(600):
        engine.frame(DT, [])
        w.frame(DT, [])
        if k % 50 == 49 or k < 3:
            assert engine.counts(1) == w.counts(1), f"frame {k}"
            assert_rows_match(engine.read_particles(1, 0), w.read_particles(1, 0), what=f"frame {k}")
    assert engine.counts(1)[0] == 0 or engine.counts(1)[0] < 5000


def test_zero_angular_velocity_keeps_rotation_exact(engine, oracle):
    sp = _idle_spawner(lifetime=RandF32.constant(5.0))
    w = oracle.OracleWorld()
    reset_both(engine, w, 1, sp)
    rows = random_rows(np.random.default_rng(1), 1000, angular=False)
    rows["lifetime"] = 5.0
    rows["age"] = np.sort(rows["age"])[::-1]
    engine.write_particles(1, 0, rows)
    w.write_particles(1, 0, rows)
    for _ in range(10):
        engine.frame(DT, [])
        w.frame(DT, [])
    assert_rows_match(engine.read_particles(1, 0), w.read_particles(1, 0))


@pytest.mark.parametrize("lifetime,frame", [(0.75, 46), (1.0, 61), (2.0, 121), (2.5, 151)])
def test_death_frames(engine, lifetime, frame):
    """sequential f32 age accumulation decides the death frame."""
    sp = ParticleSpawner(particle_settings=[ParticleSettings(lifetime=RandF32.constant(lifetime))],
                         emission_settings=[EmissionSettings(emission_pacing=EmissionPacing.OneShot(300))])
    ps, nt, es, ne = sp.pods()
    engine.spawner_reset(1, ps, nt, es, ne, True)
    for k in range(1, frame + 2):
        engine.frame(DT, [frame_input(1)])
        if k in (frame - 1, frame):
            assert engine.counts(1)[0] == (300 if k < frame else 0), k
    st = engine.status(1)
    assert st.all_empty and not st.active and st.finished


@pytest.mark.parametrize("rate", [1000.0, 15625.0, 160000.0])
def test_emission_counts_every_frame(engine, oracle, rate):
    """R5 + deaths: data.particles[i].len() equals the oracle's on every one of 130 frames."""
    sp = stress_spawner(rate=rate)
    w = oracle.OracleWorld()
    reset_both(engine, w, 9, sp)
    inp = [frame_input(9, (0.0, 0.1, 0.0))]
    for k in range(130):
        engine.frame(DT, inp)
        w.frame(DT, inp)
        assert engine.counts(9) == w.counts(9), f"frame {k}"
    assert_rows_match(engine.read_particles(9, 0), w.read_particles(9, 0))


def test_spawn_parity_all_shapes(engine, oracle):
    """R4/R8/R9 under the Philox protocol: same uniforms -> same particles, every field equal."""
    q = (0.0, math.sin(0.4), 0.0, math.cos(0.4))
    emitters = [
        EmissionSettings(emission_pacing=EmissionPacing.OneShot(1000), emission_shape=EmissionShape.Point,
                         initial_velocity=RandVec3(RandF32(1.0, 4.0), (1.0, 2.0, 0.5), 0.7),
                         initial_angular_velocity=RandVec3(RandF32(0.5, 2.0), (0.0, 0.0, 1.0), 0.3),
                         initial_rotation=q),
        EmissionSettings(emission_pacing=EmissionPacing.OneShot(777), emission_shape=EmissionShape.Sphere(2.0),
                         initial_velocity_radial=RandF32(0.5, 3.0), inherit_parent_velocity=False),
        EmissionSettings(emission_pacing=EmissionPacing.OneShot(555),
                         emission_shape=EmissionShape.Circle((0.3, 0.8, -0.2), 1.5), particle_index=1,
                         initial_velocity=RandVec3.constant((0.0, 2.0, 0.0))),
    ]
    sp = ParticleSpawner(particle_settings=[ParticleSettings(lifetime=RandF32(1.0, 2.0), initial_scale=RandF32(0.1, 0.4)),
                                            ParticleSettings(base_color=CURVES["uneven"][1])],
                         emission_settings=emitters)
    w = oracle.OracleWorld()
    reset_both(engine, w, 77, sp)
    inp = [frame_input(77, (1.0, 2.0, 3.0), (math.sin(0.3), 0.0, 0.0, math.cos(0.3)), (0.5, 0.0, -0.5), 1.5, 0.8)]
    engine.frame(DT, inp)
    w.frame(DT, inp)
    assert engine.counts(77) == w.counts(77) == [1777, 555]
    for t in (0, 1):
        assert_rows_match(engine.read_particles(77, t), w.read_particles(77, t),
                          what=f"type {t}")


def test_sparks_trajectory_c1(engine, oracle):
    """C1 examples/sparks.rs: 240 frames, counts exact every frame, every field of every row equal."""
    sp = sparks_spawner(1000.0)
    w = oracle.OracleWorld()
    reset_both(engine, w, 1, sp)
    inp = [frame_input(1, (0.0, 0.1, 0.0))]
    for k in range(240):
        engine.frame(DT, inp)
        w.frame(DT, inp)
        assert engine.counts(1) == w.counts(1), f"frame {k}"
        if k % 60 == 59:
            assert_rows_match(engine.read_particles(1, 0), w.read_particles(1, 0), what=f"frame {k}")
    bb, ob = engine.read_aabb(1), w.read_aabb(1)
    assert (np.asarray(bb[0]) == np.asarray(ob[0])).all() and (np.asarray(bb[1]) == np.asarray(ob[1])).all()


def test_stress_64_spawners_c2_reduced(engine, oracle):
    """C2 layout (64 spawners on a grid) at a reduced rate so the oracle finishes in seconds."""
    w = oracle.OracleWorld(n_threads=8)
    sp = stress_spawner(rate=1500.0)
    pos = grid_positions(64)
    inputs = []
    for i, p in enumerate(pos):
        reset_both(engine, w, 100 + i, sp)
        inputs.append(frame_input(100 + i, p))
    for k in range(150):
        engine.frame(DT, inputs)
        w.frame(DT, inputs)
    keys, types, counts = engine.counts_all()
    assert list(keys) == [100 + i for i in range(64)]
    assert [int(c) for c in counts] == [w.counts(100 + i)[0] for i in range(64)]
    assert engine.total_live() == w.total_live()
    for i in (0, 17, 63):
        assert_rows_match(engine.read_particles(100 + i, 0), w.read_particles(100 + i, 0))


def test_one_shot_bursts_c4_reduced(engine, oracle):
    """C4: a new OneShot spawner every frame, retired when finished (examples/one_shot.rs:137-141)."""
    w = oracle.OracleWorld()
    sp = one_shot_spawner(count=3000, lifetime=0.5)
    live = []
    for k in range(80):
        key = 1000 + k
        reset_both(engine, w, key, sp)
        live.append(key)
        q = (0.0, 0.0, math.sin(0.1 * k), math.cos(0.1 * k))
        inputs = [frame_input(key, (0.1 * k, 1.0, 0.0), q)]
        engine.frame(DT, inputs)
        w.frame(DT, inputs)
        for key2 in list(live):
            st, ost = engine.status(key2), w.status(key2)
            assert (st.finished, st.active, st.live_particles) == (ost.finished, ost.active, ost.live_particles)
            if st.finished:
                engine.spawner_remove(key2)
                w.spawner_remove(key2)
                live.remove(key2)
        if k % 20 == 19:
            assert engine.total_live() == w.total_live()
            assert_rows_match(engine.read_particles(live[0], 0), w.read_particles(live[0], 0))
    assert len(live) == 29  # lifetime 0.5 s: the f32 age sum reaches 0.5 on update #30


def test_random_lifetime_compaction(engine, oracle):
    """random lifetimes: deaths anywhere in the Vec; order and counts must match every frame."""
    sp = stress_spawner(rate=20000.0)
    sp.particle_settings[0].lifetime = RandF32(0.2, 1.2)
    w = oracle.OracleWorld()
    reset_both(engine, w, 3, sp)
    inp = [frame_input(3, (0.0, 0.1, 0.0))]
    for k in range(120):
        engine.frame(DT, inp)
        w.frame(DT, inp)
        assert engine.counts(3) == w.counts(3), f"frame {k}"
        if k % 40 == 39:
            assert_rows_match(engine.read_particles(3, 0), w.read_particles(3, 0), what=f"frame {k}")


def test_ring_wrap_and_growth(engine, oracle):
    """a deliberately tiny capacity hint forces ring wrap-around and several growths."""
    sp = stress_spawner(rate=6000.0, lifetime=0.4)
    sp.particle_settings[0].capacity_hint = 64
    w = oracle.OracleWorld()
    reset_both(engine, w, 3, sp)
    inp = [frame_input(3, (0.0, 0.1, 0.0))]
    for k in range(200):
        engine.frame(DT, inp)
        w.frame(DT, inp)
        if k % 10 == 0:
            assert engine.counts(3) == w.counts(3), f"frame {k}"
    engine.sync()
    assert_rows_match(engine.read_particles(3, 0), w.read_particles(3, 0))


def test_collision_single_step_prefilled(engine, oracle):
    """R3 on identical input state: bit-exact positions/velocities against the oracle's ray caster."""
    sp = _idle_spawner(lifetime=RandF32.constant(100.0), linear_drag=0.15,
                       collision_settings=ParticleCollisionSettings(0.6, 0.2, False))
    cols = [cuboid((8, 1, 8), (0, -0.5, 0)), cuboid((1, 1, 1), (0, 0.5, 0), (0.3535534, 0.3535534, 0.1464466, 0.8535534)),
            sphere(0.7, (2.0, 0.7, 0.0))]
    w = oracle.OracleWorld()
    reset_both(engine, w, 1, sp)
    engine.set_colliders(cols)
    w.set_colliders(cols)
    rng = np.random.default_rng(11)
    rows = random_rows(rng, 20000, angular=False)
    rows["position"] = rng.uniform(-3, 3, (20000, 3))
    rows["position"][:, 1] = rng.uniform(-0.2, 2.0, 20000)
    rows["velocity"] = rng.uniform(-8, 8, (20000, 3))
    rows["lifetime"] = 100.0
    rows["age"] = 1.0
    engine.write_particles(1, 0, rows)
    w.write_particles(1, 0, rows)
    for k in range(3):
        engine.frame(DT, [])
        w.frame(DT, [])
        assert_rows_match(engine.read_particles(1, 0), w.read_particles(1, 0), what=f"step {k}")


def test_collision_destroy_on_collision(engine, oracle):
    sp = _idle_spawner(lifetime=RandF32.constant(100.0),
                       collision_settings=ParticleCollisionSettings(0.6, 0.2, True))
    cols = [cuboid((8, 1, 8), (0, -0.5, 0))]
    w = oracle.OracleWorld()
    reset_both(engine, w, 1, sp)
    engine.set_colliders(cols)
    w.set_colliders(cols)
    rng = np.random.default_rng(5)
    rows = random_rows(rng, 3000, angular=False)
    rows["position"][:, 1] = rng.uniform(0.05, 3.0, 3000)
    rows["lifetime"] = 100.0
    rows["age"] = 0.0
    engine.write_particles(1, 0, rows)
    w.write_particles(1, 0, rows)
    for k in range(30):
        engine.frame(DT, [])
        w.frame(DT, [])
        assert engine.counts(1) == w.counts(1), k
    assert 0 < engine.counts(1)[0] < 3000
    assert_rows_match(engine.read_particles(1, 0), w.read_particles(1, 0))


def test_collision_scene_c5_reduced(engine, oracle):
    """C5: ring of tilted spawners over the ground slab + rotated unit cubes, 120 frames."""
    w = oracle.OracleWorld(n_threads=8)
    sp = collision_spawner(rate=600.0)
    cols = collision_scene_colliders(64)
    engine.set_colliders(cols)
    w.set_colliders(cols)
    inputs = []
    for i, (t, r) in enumerate(collision_ring(8)):
        reset_both(engine, w, 10 + i, sp)
        inputs.append(frame_input(10 + i, t, r))
    for k in range(150):
        engine.frame(DT, inputs)
        w.frame(DT, inputs)
    assert engine.total_live() == w.total_live()
    for i in range(8):
        assert_rows_match(engine.read_particles(10 + i, 0), w.read_particles(10 + i, 0), what=f"spawner {10 + i}")


def test_on_demand_and_modifier(engine, oracle):
    """OnDemand pacing drains manual_queued_count (src/core.rs:401-405); EffectModifier scales."""
    sp = ParticleSpawner(particle_settings=[ParticleSettings(lifetime=RandF32.constant(1.0), initial_scale=RandF32(0.5, 1.0))],
                         emission_settings=[EmissionSettings(emission_pacing=EmissionPacing.OnDemand,
                                                             initial_velocity=RandVec3.constant((0.0, 3.0, 0.0)))])
    w = oracle.OracleWorld()
    reset_both(engine, w, 4, sp)
    for k, q in enumerate([0, 5, 0, 120, 1, 0]):
        inp = [frame_input(4, (0, 0, 0), modifier_scale=2.0, modifier_speed=0.5, queue_particles=q)]
        engine.frame(DT, inp)
        w.frame(DT, inp)
        assert engine.counts(4) == w.counts(4)
    assert engine.counts(4)[0] == 126
    assert_rows_match(engine.read_particles(4, 0), w.read_particles(4, 0))


def test_reset_drops_particles(engine):
    """sync_spawner_data on a changed spawner drops all particles (src/core.rs:360)."""
    sp = stress_spawner(rate=5000.0)
    ps, nt, es, ne = sp.pods()
    engine.spawner_reset(1, ps, nt, es, ne, True)
    for _ in range(20):
        engine.frame(DT, [frame_input(1)])
    assert engine.counts(1)[0] > 1000
    engine.spawner_reset(1, ps, nt, es, ne, True)
    assert engine.counts(1)[0] == 0
    engine.frame(DT, [frame_input(1)])
    assert 80 <= engine.counts(1)[0] <= 90


def test_error_behaviour(engine):
    from bevy_firework_b200._native import FireworkError

    with pytest.raises(FireworkError) as e:
        engine.counts(12345, 1)
    assert e.value.code == _abi.FW_ERR_UNKNOWN_SPAWNER
    with pytest.raises(FireworkError) as e:
        engine.frame(DT, [frame_input(4242)])
    assert e.value.code == _abi.FW_ERR_UNKNOWN_SPAWNER
    sp = stress_spawner()
    ps, nt, es, ne = sp.pods()
    ps[0].base_color.times[2] = 0.1  # not increasing: the reference's UnevenCore would reject it
    with pytest.raises(FireworkError) as e:
        engine.spawner_reset(1, ps, nt, es, ne, True)
    assert e.value.code == _abi.FW_ERR_INVALID_ARGUMENT
    ps[0].base_color.n = 0
    with pytest.raises(FireworkError):
        engine.spawner_reset(1, ps, nt, es, ne, True)
    es[0].particle_index = 3
    with pytest.raises(FireworkError):
        engine.spawner_reset(1, *sp.pods()[:1], nt, es, ne, True)


def test_curves_at_the_knot_cap(engine, oracle):
    """FW_MAX_KNOTS samples per curve / gradient (even and uneven), one more is refused"""
    from bevy_firework_b200 import FireworkCurve, FireworkGradient, LinearRgba
    from bevy_firework_b200._native import FireworkError

    n = _abi.FW_MAX_KNOTS
    assert n >= 32
    rng = np.random.default_rng(9)

    def col():
        return LinearRgba(*[float(v) for v in rng.uniform(0.0, 3.0, 4)])

    ts = np.sort(rng.uniform(0.01, 0.99, n - 2))
    w = oracle.OracleWorld()
    for key, lifetime in ((1, RandF32.constant(0.6)), (2, RandF32(0.3, 0.9))):  # static kernel / compacting
        sp = stress_spawner(rate=6000.0)
        p = sp.particle_settings[0]
        p.lifetime = lifetime
        p.scale_curve = FireworkCurve.even_samples([float(v) for v in rng.uniform(0.2, 2.0, n)])
        p.base_color = FireworkGradient.uneven_samples([(0.0, col())] + [(float(t), col()) for t in ts] + [(1.0, col())])
        p.emissive_color = FireworkGradient.even_samples([col() for _ in range(n - 7)])
        if key == 2:
            p.scale_curve = FireworkCurve.uneven_samples([(0.0, 1.0)] + [(float(t), float(rng.uniform(0.1, 2.0))) for t in ts] + [(1.0, 0.0)])
            p.angular_acceleration = (0.0, 1.0, 0.0)  # the rotating (generic) kernel
        reset_both(engine, w, key, sp)
    inp = [frame_input(1, (0.0, 0.1, 0.0)), frame_input(2, (1.0, 0.1, 0.0))]
    for k in range(70):
        engine.frame(DT, inp)
        w.frame(DT, inp)
    for key in (1, 2):
        assert engine.counts(key) == w.counts(key)
        assert_rows_match(engine.read_particles(key, 0), w.read_particles(key, 0), what=f"spawner {key}")
    ps, nt, es, ne = stress_spawner().pods()
    ps[0].scale_curve.kind = _abi.FW_CURVE_EVEN
    ps[0].scale_curve.n = n + 1
    with pytest.raises(FireworkError) as e:
        engine.spawner_reset(3, ps, nt, es, ne, True)
    assert e.value.code == _abi.FW_ERR_INVALID_ARGUMENT


def test_pack_instances_device(engine):
    import torch

    sp = stress_spawner(rate=3000.0)
    keys = [1, 2, 3]
    for k in keys:
        ps, nt, es, ne = sp.pods()
        engine.spawner_reset(k, ps, nt, es, ne, True)
    inputs = [frame_input(k, (float(k), 0.0, 0.0)) for k in keys]
    for _ in range(70):
        engine.frame(DT, inputs)
    total = engine.total_live()
    buf = torch.zeros((total + 10, 16), dtype=torch.float32, device="cuda:0")
    n = engine.pack_instances_device(buf.data_ptr(), total + 10)
    assert n == total
    host = buf.cpu().numpy()[:n].view(_abi.particle_instance_dtype()).reshape(-1)
    want = np.concatenate([engine.read_instances(k, 0) for k in keys])
    assert host.tobytes() == want.tobytes()


def test_graph_replay_and_concurrent_spawn_match_plain_launches():
    """frames replayed as CUDA graphs, and frames whose spawn+first-step kernel runs concurrently
    with the update kernel, give bit-identical state to plain sequential kernel-by-kernel launches,
    including across a topology change mid-run."""
    from bevy_firework_b200._native import Engine

    results = []
    for graphs, concurrent in ((False, False), (True, True), (False, True), (True, False)):
        eng = Engine(device=0, seed=1234, graphs=graphs, concurrent_spawn=concurrent)
        sp = stress_spawner(rate=9000.0, lifetime=0.5)
        ps, nt, es, ne = sp.pods()
        eng.spawner_reset(1, ps, nt, es, ne, True)
        inp = [frame_input(1, (0.0, 0.1, 0.0))]
        for k in range(90):
            if k == 40:  # topology change: a second spawner appears
                eng.spawner_reset(2, ps, nt, es, ne, True)
                inp.append(frame_input(2, (3.0, 0.1, 0.0)))
            eng.frame(DT, inp)
        results.append((eng.read_particles(1, 0), eng.read_particles(2, 0), eng.read_aabb(1), eng.read_aabb(2), eng.counts(2)))
        eng.close()
    for r in results[1:]:
        assert r[0].tobytes() == results[0][0].tobytes()
        assert r[1].tobytes() == results[0][1].tobytes()
        assert r[2:] == results[0][2:]


def test_compaction_large_stream_many_tiles(engine, oracle):
    """one compacting stream of ~1100 tiles: the look-back needs several rounds of 256
    predecessors, deaths are scattered over the whole Vec; bit-exact state against the oracle."""
    n = 280_000
    sp = _idle_spawner(lifetime=RandF32(0.5, 3.0), linear_drag=0.3)
    w = oracle.OracleWorld()
    reset_both(engine, w, 1, sp)
    rng = np.random.default_rng(23)
    rows = random_rows(rng, n)
    # long runs of survivors and long runs of deaths, so that whole tiles publish 0 or 256
    rows["age"][50_000:120_000] = rows["lifetime"][50_000:120_000] * 0.1
    rows["age"][200_000:230_000] = rows["lifetime"][200_000:230_000] * np.float32(0.9999)
    engine.write_particles(1, 0, rows)
    w.write_particles(1, 0, rows)
    for k in range(12):
        engine.frame(DT, [])
        w.frame(DT, [])
        assert engine.counts(1) == w.counts(1), f"frame {k}"
    assert 0 < engine.counts(1)[0] < n
    assert_rows_match(engine.read_particles(1, 0), w.read_particles(1, 0))


def test_collision_bvh_many_mixed_colliders(engine, oracle):
    """the BVH broad phase against the oracle's test-every-collider loop: 600 cuboids and spheres
    of very different sizes (one of them enclosing the whole scene), overlapping boxes, a layer
    filter that hides a third of them, and a collider with a NaN transform that must never cull."""
    rng = np.random.default_rng(101)
    cols = [cuboid((40, 1, 40), (0, -0.5, 0))]
    for i in range(599):
        pos = rng.uniform(-6, 6, 3)
        pos[1] = rng.uniform(0.0, 5.0)
        layers = 1 if i % 3 else 2
        if i % 2:
            q = rng.normal(size=4)
            q /= np.linalg.norm(q)
            cols.append(cuboid(rng.uniform(0.05, 1.5, 3), pos, tuple(q), layers=layers))
        else:
            cols.append(sphere(float(rng.uniform(0.05, 0.9)), pos, layers=layers))
    cols.append(sphere(30.0, (0.0, 0.0, 0.0), layers=1))           # everything starts inside it
    cols.append(cuboid((1, 1, 1), (float("nan"), 0.0, 0.0), layers=1))
    sp = _idle_spawner(lifetime=RandF32.constant(100.0), linear_drag=0.15,
                       collision_settings=ParticleCollisionSettings(0.6, 0.2, False, 1))
    w = oracle.OracleWorld()
    reset_both(engine, w, 1, sp)
    engine.set_colliders(cols)
    w.set_colliders(cols)
    n = 30000
    rows = random_rows(rng, n, angular=False)
    rows["position"] = rng.uniform(-6, 6, (n, 3))
    rows["position"][:, 1] = rng.uniform(0.0, 5.0, n)
    rows["velocity"] = rng.uniform(-30, 30, (n, 3))      # up to ~0.9 units per frame: multi-node segments
    rows["velocity"][::7] = 0.0                          # Dir3 fallback (:758-761)
    rows["lifetime"] = 100.0
    rows["age"] = 1.0
    engine.write_particles(1, 0, rows)
    w.write_particles(1, 0, rows)
    for k in range(4):
        engine.frame(DT, [])
        w.frame(DT, [])
        assert_rows_match(engine.read_particles(1, 0), w.read_particles(1, 0), what=f"step {k}")


def test_collision_moving_colliders_every_frame(engine, oracle):
    """avian colliders move between physics steps: the collider set is re-sent before every frame
    (same count: asynchronous re-upload + BVH rebuild, frame graphs stay valid), then the count
    changes twice."""
    sp = collision_spawner(rate=30000.0)
    w = oracle.OracleWorld()
    reset_both(engine, w, 1, sp)
    inp = [frame_input(1, (0.0, 0.5, 0.0))]

    def scene(k, n):
        cols = [cuboid((12, 1, 12), (0, -0.5, 0))]
        for i in range(n):
            a = 0.7 * i + 0.05 * k
            cols.append(cuboid((0.8, 0.8, 0.8), (2.0 * math.cos(a), 1.0 + 0.5 * math.sin(0.3 * k + i), 2.0 * math.sin(a))))
            cols.append(sphere(0.4, (1.2 * math.cos(-a), 2.0 + 0.3 * math.cos(0.2 * k), 1.2 * math.sin(-a))))
        return cols

    for k in range(90):
        cols = scene(k, 12 if k < 40 else (5 if k < 60 else 20))
        engine.set_colliders(cols)
        w.set_colliders(cols)
        engine.frame(DT, inp)
        w.frame(DT, inp)
        if k % 15 == 14:
            assert engine.counts(1) == w.counts(1), f"frame {k}"
    got, want = engine.read_particles(1, 0), w.read_particles(1, 0)
    assert len(got) == len(want) > 20000
    assert_rows_match(got, want, what="moving colliders")


@pytest.mark.parametrize("kinds", ["cylinder_cone", "with_capsules"])
def test_collision_cylinder_and_cone(engine, oracle, kinds):
    """the colliders of examples/textures.rs:195,211 (circular base, cone) plus rotated and random
    ones (and capsules): bit-exact against the oracle on identical state, both broad-phase paths"""
    from bevy_firework_b200.workloads import capsule, cone, cylinder

    rng = np.random.default_rng(77)
    cols = [cylinder(4.0, 0.2, (0.0, 0.0, 0.0)), cone(0.5, 1.0, (0.0, 0.5, 0.0))]
    for i in range(60):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        pos = (float(rng.uniform(-3.5, 3.5)), float(rng.uniform(0.3, 3.0)), float(rng.uniform(-3.5, 3.5)))
        mk = cylinder if i % 2 else cone
        if kinds == "with_capsules" and i % 3 == 0:
            mk = capsule
        cols.append(mk(float(rng.uniform(0.1, 0.6)), float(rng.uniform(0.2, 1.2)), pos, tuple(q)))
    sp = _idle_spawner(lifetime=RandF32.constant(100.0), linear_drag=0.15,
                       collision_settings=ParticleCollisionSettings(0.6, 0.2, False))
    w = oracle.OracleWorld()
    reset_both(engine, w, 1, sp)
    engine.set_colliders(cols)
    w.set_colliders(cols)
    n = 30000
    rows = random_rows(rng, n, angular=False)
    rows["position"] = rng.uniform(-4.5, 4.5, (n, 3))
    rows["position"][:, 1] = rng.uniform(-0.3, 3.5, n)
    rows["velocity"] = rng.uniform(-12, 12, (n, 3))
    rows["velocity"][::5] *= 8.0            # long segments: the BVH path
    rows["velocity"][1::11, 0] = 0.0        # axis-parallel rays: the A == 0 / d.y == 0 branches
    rows["velocity"][1::11, 2] = 0.0
    rows["velocity"][2::13, 1] = 0.0
    rows["lifetime"] = 100.0
    rows["age"] = 1.0
    engine.write_particles(1, 0, rows)
    w.write_particles(1, 0, rows)
    for k in range(4):
        engine.frame(DT, [])
        w.frame(DT, [])
        assert_rows_match(engine.read_particles(1, 0), w.read_particles(1, 0), what=f"step {k}")
