"""Pin the CPU oracle: the reference's own two unit tests (G1, G2), published known-answer
vectors of the algorithms it restates, and the derived checks of SURVEY section 8c.
Runs without a GPU."""
import math

import numpy as np
import pytest

from bevy_firework_b200 import (EmissionPacing, EmissionSettings, EmissionShape, FireworkCurve,
                                FireworkGradient, LinearRgba, ParticleSettings, ParticleSpawner,
                                RandF32, RandVec3, _abi)
from bevy_firework_b200._native import frame_input
from bevy_firework_b200.workloads import sparks_spawner, stress_spawner

f32 = np.float32
DT = float(f32(1.0) / f32(60.0))


# ------------------------------------------------------------------ G1: src/core.rs:806-834
def test_g1_compute_emission_count(oracle):
    timestep = f32(0.016)
    age = f32(0.0)
    last_emission = float(np.finfo(np.float32).min)  # f32::MIN
    total = 0
    while age <= f32(3.0):
        n, last_emission = oracle.compute_emission_count(float(age), last_emission, 3.0, 0.0, 1.0, 23.0)
        total += n
        age = f32(age + timestep)
    assert total in (23, 22)  # the reference's assertion
    assert total == 22        # the value its f32 arithmetic yields


# ------------------------------------------------------------------ G2: src/curve.rs:245-258
def test_g2_curve_linear_rgba(oracle):
    g = FireworkGradient.even_samples([LinearRgba(1, 0, 0, 1), LinearRgba(0, 1, 0, 1), LinearRgba(0, 0, 1, 1)]).to_pod()
    assert oracle.sample_gradient(g, 0.0) == (1.0, 0.0, 0.0, 1.0)
    assert oracle.sample_gradient(g, 0.5) == (0.0, 1.0, 0.0, 1.0)
    assert oracle.sample_gradient(g, 1.0) == (0.0, 0.0, 1.0, 1.0)


def test_stress_gradient_values(oracle):
    """derived values of the examples/stress_test.rs:100-106 gradient."""
    g = stress_spawner().particle_settings[0].base_color.to_pod()
    cases = {0.35: (6.5, 4.0, 1.0, 1.0), 0.75: (2.0, 0.65, 0.65, 1.0), 0.95: (0.2, 0.2, 0.2, 0.5),
             1.0: (0.1, 0.1, 0.1, 0.0), 7.0: (0.1, 0.1, 0.1, 0.0), 0.0: (10.0, 7.0, 1.0, 1.0),
             -3.0: (10.0, 7.0, 1.0, 1.0)}
    for t, want in cases.items():
        got = oracle.sample_gradient(g, t)
        assert np.allclose(got, want, rtol=2e-7, atol=1e-7), (t, got)
    # exact at the knots
    assert oracle.sample_gradient(g, float(f32(0.7))) == (3.0, 1.0, 1.0, 1.0)


def test_curve_kinds(oracle):
    c = FireworkCurve.constant(3.5).to_pod()
    assert oracle.sample_curve(c, 0.3) == 3.5
    e = FireworkCurve.even_samples([1.0, 2.0]).to_pod()  # examples/one_shot.rs:100
    assert oracle.sample_curve(e, 0.0) == 1.0 and oracle.sample_curve(e, 1.0) == 2.0
    assert oracle.sample_curve(e, 0.25) == 1.25 and oracle.sample_curve(e, 2.0) == 2.0
    u = FireworkCurve.uneven_samples([(0.0, 0.0), (0.2, 1.0), (1.0, 0.5)]).to_pod()
    assert oracle.sample_curve(u, float(f32(0.2))) == 1.0
    assert abs(oracle.sample_curve(u, 0.1) - 0.5) < 1e-6
    assert abs(oracle.sample_curve(u, 0.6) - 0.75) < 1e-6
    assert oracle.sample_curve(u, -1.0) == 0.0 and oracle.sample_curve(u, 1.5) == 0.5


def test_curve_constructor_rules():
    """src/curve.rs:40-75: 0 samples is an error, 1 sample a constant, >= 2 a sample curve."""
    with pytest.raises(ValueError):
        FireworkCurve.even_samples([])
    with pytest.raises(ValueError):
        FireworkGradient.uneven_samples([])
    assert FireworkCurve.even_samples([2.0]).kind == _abi.FW_CURVE_CONSTANT
    assert FireworkCurve.uneven_samples([(0.3, 2.0)]).kind == _abi.FW_CURVE_CONSTANT
    assert FireworkGradient.even_samples([LinearRgba.WHITE, LinearRgba.BLACK]).kind == _abi.FW_CURVE_EVEN


# ------------------------------------------------------------------ Philox4x32-10 KAT
def test_philox_known_answers(oracle):
    """Random123 kat_vectors, philox4x32 10 rounds."""
    assert oracle.philox4x32_10([0, 0, 0, 0], [0, 0]) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert oracle.philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert oracle.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == (
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


def test_uniform_range_and_protocol(oracle):
    us = [oracle.uniform(0x00F12E00, 7, 0, s, d) for s in range(200) for d in range(12)]
    assert all(0.0 <= u < 1.0 for u in us)
    assert abs(np.mean(us) - 0.5) < 0.03
    # draw d of a particle = lane d&3 of block d>>2
    out = oracle.philox4x32_10([5, 0, 7, (3 << 8) | 1], [0x00F12E00, 0])
    assert oracle.uniform(0x00F12E00, 7, 3, 5, 6) == (out[2] >> 8) * 2.0 ** -24


# ------------------------------------------------------------------ Rust f32 helpers
def test_euclid_helpers(oracle):
    L = oracle.lib()
    assert L.fwo_rem_euclid(7.5, 2.0) == 1.5 and L.fwo_rem_euclid(-0.5, 2.0) == 1.5
    assert L.fwo_div_euclid(7.5, 2.0) == 3.0 and L.fwo_div_euclid(-0.5, 2.0) == -1.0
    assert L.fwo_div_euclid(-4.0, 2.0) == -2.0


# ------------------------------------------------------------------ death frames
@pytest.mark.parametrize("lifetime,frame", [(0.75, 46), (1.0, 61), (2.0, 121), (2.5, 151)])
def test_death_frames(oracle, lifetime, frame):
    w = oracle.OracleWorld()
    sp = ParticleSpawner(particle_settings=[ParticleSettings(lifetime=RandF32.constant(lifetime))],
                         emission_settings=[EmissionSettings(emission_pacing=EmissionPacing.OneShot(3))])
    ps, nt, es, ne = sp.pods()
    w.spawner_reset(1, ps, nt, es, ne, True)
    removed_at = None
    for k in range(1, 200):
        w.frame(DT, [frame_input(1)])
        if w.counts(1)[0] == 0:
            removed_at = k
            break
        assert w.counts(1)[0] == 3
    assert removed_at == frame


# ------------------------------------------------------------------ emission sequences
@pytest.mark.parametrize("rate,head,total3s", [(1000.0, [16, 17, 17, 16], 2965), (160000.0, [2666, 2667, 2667, 2666], 474668),
                                                (15625.0, [260, 260, 261, 260], 46352)])
def test_emission_sequences(oracle, rate, head, total3s):
    t, last = 0.0, 0.0
    counts = []
    for _ in range(180):
        t = oracle.lib().fwo_rem_euclid(float(f32(t) + f32(DT)), 1.0)
        n, last = oracle.compute_emission_count(t, last, 1.0, 0.0, 1.0, rate)
        counts.append(n)
    assert counts[:4] == head
    assert sum(counts) == total3s
    assert counts.count(0) == 2 or counts.count(0) == 3  # one zero-emission frame per cycle wrap


# ------------------------------------------------------------------ closed-form Euler
def test_closed_form_euler(oracle):
    """v_n = (v0 - a/k)(1-k dt)^n + a/k ; p_n = p0 + dt * sum v_i  (SURVEY section 8c)."""
    w = oracle.OracleWorld()
    sp = ParticleSpawner(particle_settings=[ParticleSettings(lifetime=RandF32.constant(10.0), linear_drag=0.1)],
                         emission_settings=[EmissionSettings(emission_pacing=EmissionPacing.OneShot(0))])
    ps, nt, es, ne = sp.pods()
    w.spawner_reset(1, ps, nt, es, ne, True)
    rows = np.zeros(1, dtype=_abi.particle_data_dtype())
    rows["position"] = (1.0, 2.0, 3.0)
    rows["velocity"] = (4.0, 9.0, -2.0)
    rows["rotation"] = (0, 0, 0, 1)
    rows["lifetime"] = 10.0
    rows["initial_scale"] = 1.0
    w.write_particles(1, 0, rows)
    n = 44
    for _ in range(n):
        w.update_only(DT)
    got = w.read_particles(1, 0)[0]
    a, k, dt = np.array([0.0, -9.81, 0.0]), 0.1, float(f32(DT))
    v0, p0 = np.array([4.0, 9.0, -2.0]), np.array([1.0, 2.0, 3.0])
    vs = [(v0 - a / k) * (1 - k * dt) ** i + a / k for i in range(n + 1)]
    p = p0 + dt * np.sum(vs[:n], axis=0)
    assert np.allclose(got["velocity"], vs[n], rtol=2e-6, atol=1e-6)
    assert np.allclose(got["position"], p, rtol=2e-6, atol=1e-6)
    assert got["age"] == pytest.approx(n * dt, rel=1e-5)


# ------------------------------------------------------------------ samplers (build-defined)
def test_spawn_statistics(oracle):
    """ranges / cone half-angle / disk radius of the spawn samplers."""
    es = sparks_spawner().emission_settings[0].to_pod()
    rng = np.random.default_rng(0)
    for _ in range(500):
        u = rng.random(3)
        p = oracle.generate_point(es, *u)
        assert abs(p[1]) < 1e-6 and math.hypot(p[0], p[2]) <= 0.3 + 1e-6   # disk in the XZ plane
        v = oracle.rand_vec3(es.initial_velocity, *rng.random(3))
        m = np.linalg.norm(v)
        assert m <= 10.0 + 1e-5
        if m > 1e-3:
            assert math.acos(min(1.0, v[1] / m)) <= math.pi / 6 + 1e-4  # cone around +Y
    sph = EmissionSettings(emission_shape=EmissionShape.Sphere(2.0)).to_pod()
    for _ in range(200):
        assert np.linalg.norm(oracle.generate_point(sph, *rng.random(3))) <= 2.0 + 1e-5
    pt = EmissionSettings().to_pod()
    assert oracle.generate_point(pt, 0.3, 0.4, 0.5) == (0.0, 0.0, 0.0)
    # RandVec3::constant
    c = RandVec3.constant((0.0, 3.0, 4.0)).to_pod()
    assert np.allclose(oracle.rand_vec3(c, 0.1, 0.2, 0.3), (0.0, 3.0, 4.0), atol=1e-6)


def test_sparks_live_count(oracle):
    """C1 literal settings: ~754 live particles (simulated 749-767)."""
    w = oracle.OracleWorld()
    sp = sparks_spawner(1000.0)
    ps, nt, es, ne = sp.pods()
    w.spawner_reset(1, ps, nt, es, ne, True)
    seen = []
    for k in range(240):
        w.frame(DT, [frame_input(1, (0.0, 0.1, 0.0))])
        if k >= 60:
            seen.append(w.counts(1)[0])
    assert 730 <= min(seen) and max(seen) <= 770
    rows = w.read_particles(1, 0)
    assert (rows["lifetime"] == f32(0.75)).all()
    assert (np.diff(rows["age"]) <= 0).all()  # Vec order = oldest first: what the ring layout relies on


# ------------------------------------------------------------------ collision pieces
def test_ray_casts(oracle):
    from bevy_firework_b200.workloads import cuboid, sphere

    ground = cuboid((8, 1, 8), (0, -0.5, 0))
    hit = oracle.cast_ray([ground], (0, 2, 0), (0, -1, 0), 10.0)
    assert hit is not None and hit[0] == pytest.approx(2.0) and hit[1] == (0.0, 1.0, 0.0)
    assert oracle.cast_ray([ground], (0, 2, 0), (0, -1, 0), 1.5) is None          # beyond max_distance
    assert oracle.cast_ray([ground], (0, 2, 0), (0, 1, 0), 10.0) is None          # pointing away
    inside = oracle.cast_ray([ground], (0, -0.25, 0), (1, 0, 0), 10.0)            # solid: distance 0, zero normal
    assert inside is not None and inside[0] == 0.0 and inside[1] == (0.0, 0.0, 0.0)
    ball = sphere(1.0, (0, 0, 0))
    hit = oracle.cast_ray([ball], (0, 3, 0), (0, -1, 0), 10.0)
    assert hit[0] == pytest.approx(2.0) and np.allclose(hit[1], (0, 1, 0))
    two = oracle.cast_ray([ground, ball], (0, 3, 0), (0, -1, 0), 10.0)
    assert two[2] == 1                                                            # closest hit wins
    assert oracle.cast_ray([ground], (0, 2, 0), (0, -1, 0), 10.0, filter_mask=2) is None  # filtered out


def test_particle_collision_bounce(oracle):
    from bevy_firework_b200.workloads import cuboid

    ground = cuboid((8, 1, 8), (0, -0.5, 0))
    cs = _abi.fw_collision_settings(1, 0.6, 0.2, 0, 0xFFFFFFFF)
    pos, vel, destroy = oracle.particle_collision([ground], cs, (0.0, 0.05, 0.0), (1.0, -6.0, 0.0), DT)
    assert not destroy
    assert vel[1] == pytest.approx(3.6, rel=1e-5)      # restitution 0.6 on the normal part
    assert 0.0 < vel[0] < 1.0                          # friction on the tangential part
    assert pos[1] > 0.0
    cs.destroy_on_collision = 1
    _, _, destroy = oracle.particle_collision([ground], cs, (0.0, 0.05, 0.0), (1.0, -6.0, 0.0), DT)
    assert destroy
    # miss: plain Euler step (src/core.rs:792-795)
    pos, vel, destroy = oracle.particle_collision([ground], cs, (0.0, 5.0, 0.0), (1.0, 2.0, 3.0), 0.5)
    assert pos == (0.5, 6.0, 1.5) and vel == (1.0, 2.0, 3.0) and not destroy


# ------------------------------------------------------------------ cylinder / cone ray casts
def test_cylinder_and_cone_ray_casts_known_answers(oracle):
    """the analytic cylinder / cone (defined by this build, parry casts them with GJK): closed-form
    hits of the two colliders of examples/textures.rs:195,211"""
    from bevy_firework_b200.workloads import cone, cylinder

    base = [cylinder(4.0, 0.2, (0.0, 0.0, 0.0))]             # circular base: radius 4, height 0.2
    hit = oracle.cast_ray(base, (1.0, 5.0, 0.0), (0.0, -1.0, 0.0), 10.0)
    assert hit is not None and abs(hit[0] - 4.9) < 1e-6 and hit[1] == (0.0, 1.0, 0.0)   # top cap
    hit = oracle.cast_ray(base, (10.0, 0.0, 0.0), (-1.0, 0.0, 0.0), 10.0)
    assert hit is not None and abs(hit[0] - 6.0) < 1e-6 and hit[1] == (1.0, 0.0, 0.0)   # side
    assert oracle.cast_ray(base, (10.0, 0.0, 0.0), (-1.0, 0.0, 0.0), 5.9) is None        # max_distance
    assert oracle.cast_ray(base, (1.0, 5.0, 0.0), (0.0, 1.0, 0.0), 100.0) is None        # pointing away
    assert oracle.cast_ray(base, (5.0, 5.0, 0.0), (0.0, -1.0, 0.0), 100.0) is None       # past the rim
    hit = oracle.cast_ray(base, (1.0, 0.05, 1.0), (0.0, 1.0, 0.0), 1.0)                  # inside, solid
    assert hit is not None and hit[0] == 0.0 and hit[1] == (0.0, 0.0, 0.0)

    pyramid = [cone(0.5, 1.0, (0.0, 0.5, 0.0))]              # apex at y = 1, base radius 0.5 at y = 0
    hit = oracle.cast_ray(pyramid, (5.0, 0.5, 0.0), (-1.0, 0.0, 0.0), 10.0)              # side at half height: r = 0.25
    assert hit is not None and abs(hit[0] - 4.75) < 1e-6
    n = np.array(hit[1])
    assert np.allclose(n, np.array([2.0, 1.0, 0.0]) / np.sqrt(5.0), atol=1e-6)           # slope 1 : 2
    hit = oracle.cast_ray(pyramid, (0.0, 5.0, 0.0), (0.0, -1.0, 0.0), 10.0)              # onto the apex
    assert hit is not None and abs(hit[0] - 4.0) < 1e-6 and hit[1] == (0.0, 1.0, 0.0)
    hit = oracle.cast_ray(pyramid, (0.2, -3.0, 0.0), (0.0, 1.0, 0.0), 10.0)              # base cap from below
    assert hit is not None and abs(hit[0] - 3.0) < 1e-6 and hit[1] == (0.0, -1.0, 0.0)
    assert oracle.cast_ray(pyramid, (0.4, 5.0, 0.0), (0.0, -1.0, 0.0), 4.1) is None      # too short to reach the side
    hit = oracle.cast_ray(pyramid, (0.4, 5.0, 0.0), (0.0, -1.0, 0.0), 10.0)              # side at r = 0.4: y = 0.2
    assert hit is not None and abs(hit[0] - 4.8) < 1e-6
    # the mirrored nappe above the apex is not part of the solid
    assert oracle.cast_ray(pyramid, (5.0, 1.5, 0.0), (-1.0, 0.0, 0.0), 10.0) is None


def test_capsule_ray_casts(oracle):
    """the analytic capsule (segment swept by a ball; parry casts it with GJK): closed-form hits, and random
    rays against a float64 sphere-tracing of the distance to the segment"""
    from bevy_firework_b200.workloads import capsule

    cap = [capsule(0.5, 2.0, (0.0, 0.0, 0.0))]                                          # segment y in [-1, 1], radius 0.5
    hit = oracle.cast_ray(cap, (5.0, 0.3, 0.0), (-1.0, 0.0, 0.0), 10.0)                 # side
    assert hit is not None and abs(hit[0] - 4.5) < 1e-6 and hit[1] == (1.0, 0.0, 0.0)
    hit = oracle.cast_ray(cap, (0.0, 5.0, 0.0), (0.0, -1.0, 0.0), 10.0)                 # top of the upper ball
    assert hit is not None and abs(hit[0] - 3.5) < 1e-6 and hit[1] == (0.0, 1.0, 0.0)
    hit = oracle.cast_ray(cap, (0.0, -5.0, 0.0), (0.0, 1.0, 0.0), 10.0)                 # bottom of the lower ball
    assert hit is not None and abs(hit[0] - 3.5) < 1e-6 and hit[1] == (0.0, -1.0, 0.0)
    hit = oracle.cast_ray(cap, (5.0, 1.3, 0.0), (-1.0, 0.0, 0.0), 10.0)                 # upper ball, 0.3 above its centre
    assert hit is not None and abs(hit[0] - (5.0 - 0.4)) < 1e-6
    assert np.allclose(hit[1], (0.8, 0.6, 0.0), atol=1e-6)
    assert oracle.cast_ray(cap, (5.0, 1.6, 0.0), (-1.0, 0.0, 0.0), 10.0) is None        # over the top
    assert oracle.cast_ray(cap, (5.0, 0.3, 0.0), (-1.0, 0.0, 0.0), 4.4) is None         # max_distance
    assert oracle.cast_ray(cap, (5.0, 0.3, 0.0), (1.0, 0.0, 0.0), 100.0) is None        # pointing away
    hit = oracle.cast_ray(cap, (0.2, 1.2, 0.1), (0.0, 1.0, 0.0), 1.0)                   # inside the upper ball, solid
    assert hit is not None and hit[0] == 0.0 and hit[1] == (0.0, 0.0, 0.0)
    hit = oracle.cast_ray(cap, (0.2, 3.0, 0.0), (0.0, -1.0, 0.0), 10.0)                 # inside the cylinder's shadow, from above
    assert hit is not None and abs(hit[0] - (3.0 - 1.0 - math.sqrt(0.25 - 0.04))) < 1e-6

    def sdf(p, r, h):  # distance to the capsule surface, float64
        q = p.copy()
        q[1] -= np.clip(q[1], -h, h)
        return np.linalg.norm(q) - r

    rng = np.random.default_rng(11)
    r, h = 0.4, 0.75
    cap = [capsule(r, 2.0 * h, (0.0, 0.0, 0.0))]
    hits = 0
    for _ in range(1500):
        o = rng.uniform(-2.5, 2.5, 3)
        d = rng.uniform(-0.6, 0.6, 3) - o
        d /= np.linalg.norm(d)
        got = oracle.cast_ray(cap, tuple(float(np.float32(c)) for c in o), tuple(float(np.float32(c)) for c in d), 20.0)
        o64 = np.array([np.float32(c) for c in o], dtype=np.float64)
        d64 = np.array([np.float32(c) for c in d], dtype=np.float64)
        if sdf(o64, r, h) <= 0.0:
            assert got is not None and got[0] == 0.0
            continue
        t, want = 0.0, None
        for _ in range(400):  # sphere tracing: never oversteps a convex solid
            s = sdf(o64 + d64 * t, r, h)
            if s < 1e-9:
                want = t
                break
            t += s
            if t > 20.0:
                break
        if want is None:
            if got is not None:  # a grazing ray may be decided either way within rounding
                assert sdf(o64 + d64 * got[0], r, h) < 1e-4
            continue
        assert got is not None and abs(got[0] - want) < 2e-5 * max(1.0, want), (o, d, got, want)
        p = o64 + d64 * want
        n = p.copy()
        n[1] -= np.clip(n[1], -h, h)
        assert np.allclose(got[1], n / np.linalg.norm(n), atol=2e-4)
        hits += 1
    assert hits > 300


def test_exact_ray_tests_never_hit_away_from_the_solid(oracle):
    """the broad phase of the kernels may skip a collider whose box the ray segment misses only if the exact
    test cannot report a hit there. Found by the random campaign (seeds 66, 69): a ray almost along a cone's
    slant made the textbook quadratic cancel to 0 / A and reported a hit at distance -0 for a cone metres
    away. Every hit of every kind must lie on the collider's bounding ball; rays along the slant included."""
    from bevy_firework_b200.workloads import capsule, cone, cuboid, cylinder, sphere

    # the ray of seed 66, frame 80
    c = cone(0.2573815882205963, 2 * 0.5490826368331909, (1.855310320854187, 0.6611395478248596, -2.5667452812194824),
             (-0.6398850679397583, 0.5684993267059326, 0.5101174712181091, 0.08447343856096268))
    assert oracle.cast_ray([c], (1.1446307, 5.8628054, -0.7957), (0.88509476, 0.12219622, -0.4490828), 0.16555828) is None

    rng = np.random.default_rng(3)
    checked = 0
    for i in range(3000):
        kind = i % 5
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        tr = rng.uniform(-2, 2, 3)
        r, hh = float(rng.uniform(0.1, 0.8)), float(rng.uniform(0.2, 1.5))
        if kind == 0:
            he = rng.uniform(0.1, 1.0, 3)
            col, bound = cuboid(tuple(2 * he), tr, tuple(q)), float(np.linalg.norm(he))
        elif kind == 1:
            col, bound = sphere(r, tr), r
        elif kind == 2:
            col, bound = cylinder(r, hh, tr, tuple(q)), math.hypot(r, hh / 2)
        elif kind == 3:
            col, bound = cone(r, hh, tr, tuple(q)), math.hypot(r, hh / 2)
        else:
            col, bound = capsule(r, hh, tr, tuple(q)), r + hh / 2
        x, y, z, w = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        for j in range(6):
            o = tr + rng.uniform(-6, 6, 3)
            if j < 3 and kind in (2, 3):  # along the slant (cone) / the axis (cylinder), up to rounding
                slope = (r / hh) if kind == 3 else 0.0
                phi = rng.uniform(0, 2 * np.pi)
                dl = np.array([slope * np.cos(phi), -1.0, slope * np.sin(phi)]) * rng.choice([-1.0, 1.0])
                d = R @ dl
            else:
                d = rng.normal(size=3)
            d = d / np.linalg.norm(d)
            md = float(rng.choice([0.2, 3.0, 50.0]))
            hit = oracle.cast_ray([col], tuple(float(np.float32(v)) for v in o), tuple(float(np.float32(v)) for v in d), md)
            checked += 1
            if hit is None:
                continue
            assert 0.0 <= hit[0] <= md * (1 + 1e-6)
            p = o + d * hit[0]
            assert np.linalg.norm(p - tr) <= bound * (1 + 1e-3) + 1e-3 or hit[0] == 0.0 and np.linalg.norm(o - tr) <= bound * (1 + 1e-3) + 1e-3, \
                (kind, o, d, md, hit)
    assert checked == 18000


def test_culled_ray_cast_equals_brute_force(oracle):
    """the oracle's test helper (conservative boxes, used by the full-size collision scenes) returns
    exactly what the brute-force loop over every collider returns: hit or not, distance, normal, index"""
    import ctypes as C

    from bevy_firework_b200 import _abi
    from bevy_firework_b200.workloads import capsule, collision_scene_colliders, cone, cylinder, sphere

    cols = list(collision_scene_colliders(64)) + [sphere(0.5, (1.0, 1.0, 1.0)), cylinder(0.6, 1.0, (-2.0, 1.0, 0.5)),
                                                  cone(0.5, 1.2, (0.5, 0.6, -2.0)),
                                                  capsule(0.3, 1.4, (2.5, 1.2, -1.0), (0.0, 0.0, 0.38268343, 0.92387953))]
    arr = (_abi.fw_collider * len(cols))(*cols)
    rng = np.random.default_rng(5)
    hits = 0
    for _ in range(4000):
        o = rng.uniform(-7, 7, 3) * (1.0, 0.4, 1.0) + (0.0, 1.0, 0.0)
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        md = float(rng.choice([0.05, 0.2, 1.0, 30.0]))
        a = oracle.cast_ray(arr, o, d, md)
        b = oracle.cast_ray(arr, o, d, md, culled=True)
        assert a == b, (o, d, md, a, b)
        hits += a is not None
    assert hits > 500
    inf = float("inf")
    assert oracle.cast_ray(arr, (0, 5, 0), (0, -1, 0), inf) == oracle.cast_ray(arr, (0, 5, 0), (0, -1, 0), inf, culled=True)


def test_cuboid_edge_hit_normal(oracle):
    """a ray that enters a cuboid exactly through an edge (two slabs at the same parameter): parry's
    clip_aabb_line reports -dir.normalize(); a zero normal would turn the bounce of
    src/core.rs:778-784 into NaNs"""
    from bevy_firework_b200.workloads import cuboid

    cube = cuboid((1.0, 1.0, 1.0), (0.0, 0.0, 0.0))
    d = np.array([1.0, 1.0, 0.0], dtype=np.float32) / np.sqrt(np.float32(2.0))
    hit = oracle.cast_ray([cube], (-1.5, -1.5, 0.0), tuple(d), 10.0)
    assert hit is not None and hit[0] == pytest.approx(math.sqrt(2.0), rel=1e-6)
    assert np.allclose(hit[1], -d, atol=1e-7)
    cs = _abi.fw_collision_settings(1, 0.5, 0.1, 0, 0xFFFFFFFF)
    pos, vel, destroyed = oracle.particle_collision([cube], cs, (-0.55, -0.55, 0.0), (6.0, 6.0, 0.0), 1.0 / 60.0)
    assert np.isfinite(pos).all() and np.isfinite(vel).all() and not destroyed
    assert vel[0] < 0 and vel[1] < 0  # bounced straight back along the diagonal
