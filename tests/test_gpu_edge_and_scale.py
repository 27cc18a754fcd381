"""Edge cases the reference handles implicitly (empty / degenerate inputs, dt = 0, bursts, calls
from several threads) and size-independent properties at BASELINE.json's full sizes (C3: 10 M
particles in 512 streams), where the oracle is too slow to replay everything."""
import threading

import numpy as np
import pytest

from bevy_firework_b200 import (EmissionPacing, EmissionSettings, ParticleSettings, ParticleSpawner, RandF32,
                                RandVec3, _abi)
from bevy_firework_b200._native import Engine, FireworkError, frame_input
from bevy_firework_b200.workloads import grid_positions, stress_spawner
from _parity import assert_rows_match, reset_both

pytestmark = pytest.mark.gpu
f32 = np.float32
DT = float(f32(1.0) / f32(60.0))


def test_empty_context_and_degenerate_spawners(engine, oracle):
    engine.frame(DT, [])                                  # no spawners at all
    engine.sync()
    assert engine.total_live() == 0
    keys, types, counts = engine.counts_all()
    assert len(keys) == 0
    # a spawner with particle types but no emitters, and one whose emitter never fires
    sp0 = ParticleSpawner(particle_settings=[ParticleSettings()], emission_settings=[])
    sp1 = ParticleSpawner(emission_settings=[EmissionSettings(emission_pacing=EmissionPacing.rate(0.0))])
    w = oracle.OracleWorld()
    reset_both(engine, w, 1, sp0)
    reset_both(engine, w, 2, sp1)
    for _ in range(5):
        engine.frame(DT, [frame_input(1), frame_input(2)])
        w.frame(DT, [frame_input(1), frame_input(2)])
    assert engine.counts(1) == w.counts(1) == [0] and engine.counts(2) == w.counts(2) == [0]
    assert engine.read_aabb(1) is None
    st, ost = engine.status(1), w.status(1)
    assert (st.active, st.all_empty, st.finished) == (ost.active, ost.all_empty, ost.finished)
    assert len(engine.read_particles(2, 0)) == 0 and len(engine.read_instances(2, 0)) == 0


def test_zero_dt_and_varying_dt(engine, oracle):
    """dt is whatever Time::delta_secs() says: 0 (paused virtual time) and uneven steps."""
    sp = stress_spawner(rate=4000.0, lifetime=0.3)
    w = oracle.OracleWorld()
    reset_both(engine, w, 1, sp)
    rng = np.random.default_rng(0)
    dts = [DT, 0.0, 0.0, 0.004, 0.05, DT, 0.0] + [float(f32(x)) for x in rng.uniform(0.001, 0.04, 60)]
    for k, dt in enumerate(dts):
        engine.frame(dt, [frame_input(1, (0.0, 0.1, 0.0))])
        w.frame(dt, [frame_input(1, (0.0, 0.1, 0.0))])
        assert engine.counts(1) == w.counts(1), f"frame {k} dt {dt}"
    assert_rows_match(engine.read_particles(1, 0), w.read_particles(1, 0))


def test_large_one_shot_burst(engine):
    """EmissionPacing::OneShot(5_000_000) into one stream: ring sized on demand, every particle
    spawned exactly once, all identical in age/lifetime, serials all distinct."""
    n = 5_000_000
    sp = ParticleSpawner(particle_settings=[ParticleSettings(lifetime=RandF32.constant(1.0), initial_scale=RandF32(0.0, 1.0))],
                         emission_settings=[EmissionSettings(emission_pacing=EmissionPacing.OneShot(n),
                                                             initial_velocity=RandVec3(RandF32(0.0, 1.0), (0.0, 1.0, 0.0), 0.5))])
    ps, nt, es, ne = sp.pods()
    engine.spawner_reset(1, ps, nt, es, ne, True)
    engine.frame(DT, [frame_input(1)])
    assert engine.counts(1) == [n]
    rows = engine.read_particles(1, 0)
    assert (rows["age"] == f32(DT)).all() and (rows["lifetime"] == 1.0).all()
    # initial_scale = u * 1: a 24-bit uniform per particle from distinct Philox counters
    assert 0.499 < rows["initial_scale"].mean() < 0.501
    assert len(np.unique(rows["initial_scale"])) > 0.25 * (1 << 24) * (1 - np.exp(-n / (1 << 24)))
    for _ in range(61):
        engine.frame(DT, [frame_input(1)])
    assert engine.counts(1) == [0] and engine.status(1).finished


def test_calls_from_different_threads(engine, oracle):
    """Bevy runs the systems on arbitrary worker threads: same context, one call at a time."""
    sp = stress_spawner(rate=3000.0)
    w = oracle.OracleWorld()
    reset_both(engine, w, 1, sp)
    errors = []

    def work(k):
        try:
            for _ in range(10):
                engine.frame(DT, [frame_input(1, (0.0, 0.1, 0.0))])
            engine.counts(1)
        except Exception as e:  # pragma: no cover
            errors.append(e)

    for k in range(6):
        t = threading.Thread(target=work, args=(k,))
        t.start()
        t.join()
    for _ in range(60):
        w.frame(DT, [frame_input(1, (0.0, 0.1, 0.0))])
    assert not errors
    assert engine.counts(1) == w.counts(1)


def test_two_contexts_are_independent():
    a, b = Engine(device=0, seed=1), Engine(device=0, seed=1)
    sp = stress_spawner(rate=2000.0)
    ps, nt, es, ne = sp.pods()
    for e in (a, b):
        e.spawner_reset(5, ps, nt, es, ne, True)
    for k in range(30):
        a.frame(DT, [frame_input(5)])
        if k % 2 == 0:
            b.frame(DT, [frame_input(5)])
    assert a.counts(5)[0] > b.counts(5)[0] > 0
    for _ in range(15):
        b.frame(DT, [frame_input(5)])
    assert a.read_particles(5, 0).tobytes() == b.read_particles(5, 0).tobytes()  # same seed, same history
    a.close()
    b.close()


def test_unknown_and_removed_spawners(engine):
    sp = stress_spawner(rate=100.0)
    ps, nt, es, ne = sp.pods()
    engine.spawner_reset(9, ps, nt, es, ne, True)
    engine.frame(DT, [frame_input(9)])
    engine.spawner_remove(9)
    for call in (lambda: engine.counts(9, 1), lambda: engine.status(9), lambda: engine.read_particles(9, 0),
                 lambda: engine.spawner_remove(9), lambda: engine.frame(DT, [frame_input(9)])):
        with pytest.raises(FireworkError) as e:
            call()
        assert e.value.code == _abi.FW_ERR_UNKNOWN_SPAWNER
    engine.frame(DT, [])  # the context keeps working


def test_c3_full_size_properties(oracle):
    """BASELINE config 3 at full size (512 spawners x rate 19531, ~10 M particles): counts per
    stream must equal the reference's emission arithmetic exactly; ages/lifetimes/colours are
    functions of the spawn frame only; the per-spawner AABBs must bound every instance row."""
    import torch

    n_sp, rate, frames = 512, 19531.0, 75
    eng = Engine(device=0, seed=0x00F12E00)
    sp = stress_spawner(rate=rate)
    pos = grid_positions(n_sp)
    ps, nt, es, ne = sp.pods()
    inputs = []
    for i in range(n_sp):
        eng.spawner_reset(1 + i, ps, nt, es, ne, True)
        inputs.append(frame_input(1 + i, pos[i]))
    for _ in range(frames):
        eng.frame(DT, inputs)
    # expected live count: particles emitted in the last 60 frames (lifetime 1 s dies on update #61)
    t = last = 0.0
    emitted = []
    for _ in range(frames):
        t = oracle.lib().fwo_rem_euclid(float(f32(t) + f32(DT)), 1.0)
        n, last = oracle.compute_emission_count(t, last, 1.0, 0.0, 1.0, rate)
        emitted.append(n)
    want = sum(emitted[-60:])
    keys, types, counts = eng.counts_all()
    assert len(counts) == n_sp and (counts == want).all()
    assert eng.total_live() == want * n_sp
    # one stream in detail: ages are the f32 partial sums of dt, grouped by spawn frame, oldest first
    rows = eng.read_particles(1 + 300, 0)
    ages, acc = [], f32(0.0)
    for _ in range(60):
        acc = f32(acc + f32(DT))
        ages.append(acc)
    want_ages = np.concatenate([np.full(emitted[-60 + j], ages[59 - j], dtype=np.float32) for j in range(60)])
    assert (rows["age"] == want_ages).all() and (rows["lifetime"] == 1.0).all()
    # AABB of every spawner bounds its instance rows; extract = concatenation of the streams
    cap = want * n_sp + 1024
    host = torch.empty((cap, 16), dtype=torch.float32, pin_memory=True)
    n_rows = eng.extract_instances(host.data_ptr(), cap)
    assert n_rows == want * n_sp
    inst = host.numpy()[:n_rows].view(_abi.particle_instance_dtype()).reshape(n_sp, want)
    for i in (0, 137, 511):
        lo, hi = eng.read_aabb(1 + i)
        p, s = inst[i]["position"], inst[i]["scale"][:, None]
        assert ((p - s).min(axis=0) == np.array(lo, dtype=np.float32)).all()
        assert ((p + s).max(axis=0) == np.array(hi, dtype=np.float32)).all()
    assert (inst[300]["position"] == rows["position"]).all()
    eng.close()


def _fuzz_seeds():
    """seeds 1..6 always; FW_FUZZ_SEEDS=a-b adds a longer campaign (profiles/r2: 7-160 run once on the GPU box)"""
    import os

    seeds = [1, 2, 3, 4, 5, 6]
    extra = os.environ.get("FW_FUZZ_SEEDS", "")
    if "-" in extra:
        a, b = extra.split("-")
        seeds += [s for s in range(int(a), int(b) + 1) if s not in seeds]
    return seeds


@pytest.mark.parametrize("seed", _fuzz_seeds())
def test_randomized_mixed_scene(engine, oracle, seed):
    """a seeded random scene replayed on both sides: spawners of every update variant in one
    context (FIFO, compacting, colliding, colliding + destroy), random shapes / curves / rates,
    uneven dt, spawners reset and removed on the way. Counts every frame; at the end every field of
    every row of every stream equal."""
    from bevy_firework_b200 import (EmissionShape, FireworkCurve, FireworkGradient, LinearRgba,
                                    ParticleCollisionSettings)
    from bevy_firework_b200.workloads import cuboid, sphere

    rng = np.random.default_rng(1000 + seed)

    def rand_curve():
        k = rng.integers(0, 3)
        if k == 0:
            return FireworkCurve.constant(float(rng.uniform(0.5, 2.0)))
        if k == 1:
            return FireworkCurve.even_samples([float(v) for v in rng.uniform(0.2, 2.0, rng.integers(2, 6))])
        ts = np.sort(rng.uniform(0.05, 0.95, rng.integers(1, 5)))
        return FireworkCurve.uneven_samples([(0.0, 1.0)] + [(float(t), float(rng.uniform(0.1, 2.0))) for t in ts] + [(1.0, 0.0)])

    def rand_gradient():
        def col():
            return LinearRgba(*[float(v) for v in rng.uniform(0.0, 4.0, 4)])
        k = rng.integers(0, 3)
        if k == 0:
            return FireworkGradient.constant(col())
        if k == 1:
            return FireworkGradient.even_samples([col() for _ in range(rng.integers(2, 7))])
        ts = np.sort(rng.uniform(0.05, 0.95, rng.integers(1, 6)))
        return FireworkGradient.uneven_samples([(0.0, col())] + [(float(t), col()) for t in ts] + [(1.0, col())])

    def rand_spawner(kind):
        life = float(rng.uniform(0.2, 0.9))
        lifetime = RandF32.constant(life) if kind in ("fifo", "collide") else RandF32(0.5 * life, 1.5 * life)
        collision = None
        if kind in ("collide", "collide_destroy"):
            collision = ParticleCollisionSettings(float(rng.uniform(0.2, 0.9)), float(rng.uniform(0.0, 0.5)),
                                                  kind == "collide_destroy")
        shape = [EmissionShape.Point, EmissionShape.Sphere(float(rng.uniform(0.1, 0.6))),
                 EmissionShape.Circle((0.0, 1.0, 0.0), float(rng.uniform(0.1, 0.6)))][rng.integers(0, 3)]
        return ParticleSpawner(
            particle_settings=[ParticleSettings(
                lifetime=lifetime, initial_scale=RandF32(0.02, 0.2), scale_curve=rand_curve(),
                base_color=rand_gradient(), emissive_color=rand_gradient(),
                linear_drag=float(rng.uniform(0.0, 0.5)), angular_drag=float(rng.uniform(0.0, 0.5)),
                angular_acceleration=(0.0, float(rng.uniform(-2, 2)), 0.0), collision_settings=collision,
                capacity_hint=int(rng.choice([0, 64, 4096])))],
            emission_settings=[EmissionSettings(
                emission_pacing=EmissionPacing.rate(float(rng.uniform(500.0, 40000.0))), emission_shape=shape,
                initial_velocity=RandVec3(RandF32(1.0, 9.0), (0.0, 1.0, 0.0), float(rng.uniform(0.0, 1.0))),
                initial_velocity_radial=RandF32(0.0, float(rng.uniform(0.0, 2.0))),
                initial_angular_velocity=RandVec3(RandF32(0.0, 3.0), (0.0, 1.0, 0.0), 0.5))])

    cols = [cuboid((30, 1, 30), (0, -0.5, 0))]
    for _ in range(40):
        p = (float(rng.uniform(-5, 5)), float(rng.uniform(0.3, 4.0)), float(rng.uniform(-5, 5)))
        cols.append(cuboid(tuple(rng.uniform(0.3, 1.5, 3)), p) if rng.integers(0, 2) else sphere(float(rng.uniform(0.2, 0.8)), p))
    w = oracle.OracleWorld()
    engine.set_colliders(cols)
    w.set_colliders(cols)
    kinds = ["fifo", "compact", "collide", "collide_destroy"]
    spawners, inputs = {}, {}
    for key in range(1, 13):
        kind = kinds[(key + seed) % 4]
        spawners[key] = (kind, rand_spawner(kind))
        reset_both(engine, w, key, spawners[key][1])
        inputs[key] = frame_input(key, (float(rng.uniform(-4, 4)), float(rng.uniform(0.3, 2.0)), float(rng.uniform(-4, 4))))
    for k in range(110):
        dt = float(f32(rng.choice([1 / 60, 1 / 60, 1 / 144, 1 / 30, 0.0])))
        if k in (35, 70):                      # Changed<ParticleSpawner>: reset drops the particles
            key = int(rng.choice(list(spawners)))
            reset_both(engine, w, key, spawners[key][1])
        if k == 50:                            # entity despawned
            key = int(rng.choice(list(spawners)))
            engine.spawner_remove(key)
            w.spawner_remove(key)
            del spawners[key], inputs[key]
        inp = list(inputs.values())
        engine.frame(dt, inp)
        w.frame(dt, inp)
        for key in spawners:
            assert engine.counts(key) == w.counts(key), f"frame {k} spawner {key} ({spawners[key][0]})"
    for key, (kind, _) in spawners.items():
        assert_rows_match(engine.read_particles(key, 0), w.read_particles(key, 0), what=f"spawner {key} ({kind})")


def test_failed_frame_leaves_pacing_untouched():
    """fw_frame advances emission clocks while it plans the frame; when a later step fails (here: a
    OneShot that would exceed 2^32 particles in one stream) every emitter and spawner is put back, so
    the failed call changed nothing (non-zero status, state untouched)"""
    from bevy_firework_b200._native import Engine, FireworkError

    a = stress_spawner(rate=1700.0)
    huge = ParticleSpawner(particle_settings=[ParticleSettings(lifetime=RandF32.constant(1.0))],
                           emission_settings=[EmissionSettings(emission_pacing=EmissionPacing.OneShot(5_000_000_000))])
    inp = [frame_input(1, (0.0, 0.1, 0.0))]

    def run(with_failure):
        eng = Engine(device=0, seed=0x00F12E00)
        ps, nt, es, ne = a.pods()
        eng.spawner_reset(1, ps, nt, es, ne, True)
        counts = []
        for k in range(12):
            if with_failure and k == 4:
                hps, hnt, hes, hne = huge.pods()
                eng.spawner_reset(2, hps, hnt, hes, hne, True)
                for _ in range(2):  # fails the same way twice: the OneShot emitter was re-enabled too
                    with pytest.raises(FireworkError) as e:
                        eng.frame(DT, inp + [frame_input(2, (0.0, 0.0, 0.0))])
                    assert e.value.code == _abi.FW_ERR_OUT_OF_MEMORY
                assert eng.status(2).active == 1
                eng.spawner_remove(2)
            eng.frame(DT, inp)
            counts.append(eng.counts(1)[0])
        rows = eng.read_particles(1, 0)
        eng.close()
        return counts, rows

    c0, r0 = run(False)
    c1, r1 = run(True)
    assert c0 == c1
    assert r0.tobytes() == r1.tobytes()


def test_many_slow_emitters_one_command_per_particle(engine, oracle):
    """600 spawners at 45 particles/s: most spawn commands of a frame carry one particle, so a 256-particle
    chunk of the spawn kernel spans hundreds of commands (each thread finds its own in shared memory);
    two of the spawners fire fast, so chunks with one long command and many short ones exist too."""
    w = oracle.OracleWorld(n_threads=4)
    slow, fast = stress_spawner(rate=45.0, lifetime=0.8), stress_spawner(rate=40000.0, lifetime=0.2)
    pos = grid_positions(600)
    for i in range(600):
        reset_both(engine, w, 1 + i, fast if i in (7, 311) else slow)
    inputs = [frame_input(1 + i, p) for i, p in enumerate(pos)]
    for k in range(75):
        engine.frame(DT, inputs)
        w.frame(DT, inputs)
    for i in range(600):
        assert engine.counts(1 + i) == w.counts(1 + i), i
    for i in (0, 7, 8, 255, 256, 311, 599):
        assert_rows_match(engine.read_particles(1 + i, 0), w.read_particles(1 + i, 0), f"spawner {i}")


def test_long_run_ring_wraparound(engine, oracle):
    """thousands of frames on small rings: the FIFO heads wrap their blocks hundreds of times, the
    compacting streams flip their halves every frame, the emission clocks go through thousands of
    cycles, the look-back epochs advance; counts every 25 frames, every row at the end.
    FW_LONG_FRAMES overrides the length (profiles/r2: 20000 once)."""
    import os

    from bevy_firework_b200 import ParticleCollisionSettings
    from bevy_firework_b200.workloads import cuboid

    frames = int(os.environ.get("FW_LONG_FRAMES", "3000"))
    w = oracle.OracleWorld()
    cols = [cuboid((30, 1, 30), (0, -0.5, 0)), cuboid((30, 0.5, 30), (0, 2.5, 0.0))]  # floor, and a ceiling the sparks reach
    engine.set_colliders(cols)
    w.set_colliders(cols)
    kinds = {
        1: stress_spawner(rate=700.0, lifetime=0.25),                                   # FIFO, ~175 live, capacity_hint below
        2: stress_spawner(rate=900.0, lifetime=0.3),
        3: stress_spawner(rate=500.0, lifetime=0.4),
        4: stress_spawner(rate=650.0, lifetime=0.35),
    }
    kinds[1].particle_settings[0].capacity_hint = 64
    kinds[2].particle_settings[0].lifetime = RandF32(0.1, 0.5)                         # compacting
    kinds[3].particle_settings[0].collision_settings = ParticleCollisionSettings(0.5, 0.2, False)  # FIFO + sweep
    kinds[4].particle_settings[0].collision_settings = ParticleCollisionSettings(0.5, 0.2, True)   # look-back
    kinds[4].particle_settings[0].lifetime = RandF32(0.2, 0.5)
    for key, sp in kinds.items():
        reset_both(engine, w, key, sp)
    inputs = [frame_input(key, (0.3 * key, 1.5, 0.0)) for key in kinds]
    for k in range(frames):
        engine.frame(DT, inputs)
        w.frame(DT, inputs)
        if k % 25 == 24:
            for key in kinds:
                assert engine.counts(key) == w.counts(key), f"frame {k} spawner {key}"
    for key in kinds:
        assert_rows_match(engine.read_particles(key, 0), w.read_particles(key, 0), what=f"spawner {key}")
    assert (w.read_particles(3, 0)["velocity"][:, 1] < 0).any()      # some bounced off the ceiling
    assert len(w.read_particles(4, 0)) < len(w.read_particles(2, 0))  # some died on it


def test_profile_counters_and_events_while_toggling_profiling(engine, oracle):
    """fw_set_profiling / fw_profile_sum / fw_event_record: the counters are exact (particles that
    entered the update, particles spawned), timed frames run kernel by kernel and untimed ones as
    graphs -- switching between the two mid-run must not change a bit of the state."""
    sp = stress_spawner(rate=9000.0, lifetime=0.5)
    w = oracle.OracleWorld()
    reset_both(engine, w, 1, sp)
    inp = [frame_input(1, (0.0, 0.1, 0.0))]
    assert engine.stream_handle != 0
    engine.profile_reset()
    engine.event_record(0)
    entered = 0
    for k in range(120):
        if k % 20 == 0:
            engine.set_profiling((k // 20) % 2 == 1)  # 20 frames untimed (graphs once warm), 20 timed, ...
        engine.frame(DT, inp)
        w.spawn_only(DT, inp)      # spawn_particles (src/core.rs:302-330) ...
        entered += w.total_live()  # ... every particle present then enters update_particles (:577-670)
        w.update_only(DT)
    engine.event_record(1)
    engine.sync()
    assert engine.event_elapsed_ms(0, 1) > 0.0
    p, n = engine.profile_sum()
    assert n == 120 and p.kernel_launches >= 120
    rows_g, rows_w = engine.read_particles(1, 0), w.read_particles(1, 0)
    assert_rows_match(rows_g, rows_w)
    # exact counters: spawned = what the oracle's emission arithmetic spawned over the run
    t = last = 0.0
    total = 0
    for _ in range(120):
        t = oracle.lib().fwo_rem_euclid(float(f32(t) + f32(DT)), 1.0)
        c, last = oracle.compute_emission_count(t, last, 1.0, 0.0, 1.0, 9000.0)
        total += c
    assert p.particles_spawned == total and p.particles_updated == entered
    assert p.timed_frames == 60 and p.update_ms > 0.0
