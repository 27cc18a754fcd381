"""Known answers produced BY the reference's own dependencies (rust/gen_golden.rs) against the oracle.

The third-party arithmetic of the path -- glam's quaternion / vector formulas, bevy_utilitarian's
PitchYaw and RandVec3, bevy_math's curve cores, bevy_color's Mix, parry's ray casts -- lives in crates
that are absent from /root/reference, and the build image has no Rust toolchain, so the oracle's
versions are restatements ("PARITY UNPINNED"). This test closes those rows where
a toolchain exists: run rust/gen_golden.rs inside a checkout of the reference, drop the JSON at
tests/golden/reference_vectors.json, and every section below is compared with the oracle (the CUDA
kernels are bit-equal to the oracle: tests/test_gpu_*.py). Until the file exists the test SKIPS and
says so -- it does not pass vacuously.

Comparison: bit-equal, except where the crate may take an SSE2 path whose last bit differs from the
scalar formula (glam on x86-64) -- those sections allow 1 ulp and count how many records use it.
"""
import json
import math
import os

import numpy as np
import pytest

from bevy_firework_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "golden", "reference_vectors.json")


def _f(bits):
    return np.array(bits, dtype=np.uint32).view(np.float32)


def _ulps(a, b):
    a = np.asarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a)
    b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return np.abs(a - b)


@pytest.fixture(scope="module")
def vectors():
    if not os.path.exists(PATH):
        pytest.skip("tests/golden/reference_vectors.json is absent: the third-party arithmetic stays PARITY UNPINNED "
                    "(generate it with rust/gen_golden.rs where a Rust toolchain exists)")
    with open(PATH) as f:
        return json.load(f)


def _check(records, fn, max_ulp, what):
    worst, off = 0, 0
    for r in records:
        got = np.atleast_1d(np.asarray(fn(r["in"]), dtype=np.float32))
        want = np.atleast_1d(_f(r["out"]))
        u = int(_ulps(got, want).max())
        worst = max(worst, u)
        off += u != 0
    assert worst <= max_ulp, f"{what}: worst difference {worst} ulp"
    return off


def test_generator_is_committed():
    """the pipeline exists even while the vectors do not"""
    src = open(os.path.join(os.path.dirname(HERE), "rust", "gen_golden.rs")).read()
    for section in ("glam_from_scaled_axis", "glam_from_rotation_arc", "glam_mul_quat", "glam_mul_vec3", "glam_normalize_or_zero",
                    "glam_project_onto", "glam_reject_from", "glam_from_rotation_y", "pitch_yaw", "rand_vec3", "rand_f32",
                    "curve_f32", "curve_rgba", "cast_ray", "sin_cos"):
        assert f'"{section}"' in src, section


def test_glam(vectors, oracle):
    L = oracle.lib()
    import ctypes as C

    def call(name, *arrs, n_out):
        out = (C.c_float * n_out)()
        args = [(C.c_float * len(a))(*[float(x) for x in a]) for a in arrs]
        getattr(L, name)(*args, out)
        return list(out)

    _check(vectors["glam_from_scaled_axis"], lambda i: call("fwo_quat_from_scaled_axis", _f(i), n_out=4), 1, "from_scaled_axis")
    _check(vectors["glam_mul_quat"], lambda i: call("fwo_quat_mul", _f(i[0]), _f(i[1]), n_out=4), 1, "mul_quat")
    _check(vectors["glam_from_rotation_arc"], lambda i: call("fwo_quat_from_rotation_arc", _f(i[0]), _f(i[1]), n_out=4), 1, "from_rotation_arc")
    _check(vectors["glam_mul_vec3"], lambda i: call("fwo_quat_mul_vec3", _f(i[0]), _f(i[1]), n_out=3), 1, "mul_vec3")
    _check(vectors["glam_normalize_or_zero"], lambda i: call("fwo_vec3_normalize_or_zero", _f(i), n_out=3), 1, "normalize_or_zero")
    _check(vectors["glam_project_onto"], lambda i: call("fwo_vec3_project_onto", _f(i[0]), _f(i[1]), n_out=3), 1, "project_onto")
    _check(vectors["glam_reject_from"], lambda i: call("fwo_vec3_reject_from", _f(i[0]), _f(i[1]), n_out=3), 1, "reject_from")


def test_pitch_yaw_and_rotation_y(vectors, oracle):
    e = _abi.fw_emission_settings()
    e.shape_kind = _abi.FW_SHAPE_SPHERE
    e.shape_radius = 1.0

    def unit(i):
        u, v = _f(i)
        # generate_point(Sphere(1)) with u0 = u / 2pi, u1 = v / pi, r = 1 is PitchYaw(u, v).to_unit_vec() up to
        # the two multiplications, so compare through the oracle's direct export instead
        import ctypes as C

        out = (C.c_float * 3)()
        oracle.lib().fwo_pitch_yaw_to_unit_vec(C.c_float(float(u)), C.c_float(float(v)), out)
        return list(out)

    # the platform's sin/cos and include/fw_sincos.h may differ in the last bit of an operand: 2 ulp
    _check(vectors["pitch_yaw"], unit, 2, "PitchYaw::to_unit_vec")


def test_rand_vec3_family(vectors):
    """RandVec3::generate is unseedable: recover (polar angle, magnitude) from the samples and check
    the family the build assumed -- polar angle uniform on [0, spread], magnitude uniform on [min, max]"""
    for rec in vectors["rand_vec3"]:
        d = _f(rec["in"]["direction"]).astype(np.float64)
        d /= np.linalg.norm(d)
        spread = float(_f([rec["in"]["spread"]])[0])
        lo, hi = float(_f([rec["in"]["min"]])[0]), float(_f([rec["in"]["max"]])[0])
        v = np.array([_f(s) for s in rec["out"]], dtype=np.float64)
        m = np.linalg.norm(v, axis=1)
        assert (m >= lo - 1e-5).all() and (m <= hi + 1e-5).all()
        if hi > lo:
            assert abs(m.mean() - 0.5 * (lo + hi)) < 0.02 * (hi - lo)
        ok = m > 0
        ang = np.arccos(np.clip((v[ok] @ d) / m[ok], -1.0, 1.0))
        assert (ang <= spread + 1e-4).all()
        if spread > 0:
            # uniform polar angle: mean spread / 2 (a direction uniform on the cap would give a larger mean)
            assert abs(ang.mean() - 0.5 * spread) < 0.03 * spread, (ang.mean(), spread)


def test_rand_f32_family(vectors):
    for rec in vectors["rand_f32"]:
        lo, hi = _f(rec["in"])
        x = _f(rec["out"]).astype(np.float64)
        assert (x >= lo).all() and (x <= hi).all()
        if hi > lo:
            assert abs(x.mean() - 0.5 * (float(lo) + float(hi))) < 0.02 * (float(hi) - float(lo))


def test_curves(vectors, oracle):
    from bevy_firework_b200 import FireworkCurve, FireworkGradient, LinearRgba

    def f32_curve(i):
        t = float(_f([i["t"]])[0])
        if i["kind"] == "even":
            c = FireworkCurve.even_samples([float(x) for x in _f(i["values"])])
        elif i["kind"] == "uneven":
            c = FireworkCurve.uneven_samples([(float(_f([a])[0]), float(_f([v])[0])) for a, v in i["knots"]])
        else:
            c = FireworkCurve.constant(float(_f([i["value"]])[0]))
        return oracle.sample_curve(c.to_pod(), t)

    def rgba_curve(i):
        t = float(_f([i["t"]])[0])
        if i["kind"] == "even":
            g = FireworkGradient.even_samples([LinearRgba(*[float(x) for x in _f(c)]) for c in i["values"]])
        else:
            g = FireworkGradient.uneven_samples([(float(_f([a])[0]), LinearRgba(*[float(x) for x in _f(c)])) for a, c in i["knots"]])
        return oracle.sample_gradient(g.to_pod(), t)

    assert _check(vectors["curve_f32"], f32_curve, 0, "FireworkCurve<f32>::sample_clamped") == 0
    assert _check(vectors["curve_rgba"], rgba_curve, 0, "FireworkGradient::sample_clamped") == 0


def test_cast_ray(vectors, oracle):
    kinds = {"cuboid": _abi.FW_COLLIDER_CUBOID, "sphere": _abi.FW_COLLIDER_SPHERE, "cylinder": _abi.FW_COLLIDER_CYLINDER,
             "cone": _abi.FW_COLLIDER_CONE, "capsule": _abi.FW_COLLIDER_CAPSULE}
    miss = far = 0
    for r in vectors["cast_ray"]:
        i = r["in"]
        c = _abi.fw_collider()
        c.kind = kinds[i["shape"]]
        c.layers = 1
        c.half_extents[:] = [float(x) for x in _f(i["half_extents"])]
        c.translation[:] = [float(x) for x in _f(i["translation"])]
        c.rotation[:] = [float(x) for x in _f(i["rotation"])]
        got = oracle.cast_ray([c], [float(x) for x in _f(i["origin"])], [float(x) for x in _f(i["direction"])],
                              float(_f([i["max_distance"]])[0]))
        want = r["out"]
        if (got is None) != (want is None):
            miss += 1  # a grazing ray may be decided differently by GJK (cylinder / cone) and the analytic solid
            continue
        if want is None:
            continue
        d = float(_f([want["distance"]])[0])
        n = _f(want["normal"])
        if not (math.isclose(got[0], d, rel_tol=1e-5, abs_tol=1e-5) and np.allclose(got[1], n, atol=1e-4)):
            far += 1
    n = len(vectors["cast_ray"])
    assert miss <= 0.002 * n and far <= 0.002 * n, (miss, far, n)


def test_platform_sin_cos_against_fw_sincos(vectors, oracle):
    x = _f([r["in"] for r in vectors["sin_cos"]])
    want = np.array([_f(r["out"]) for r in vectors["sin_cos"]])
    s, c = oracle.sincosf(x)
    assert _ulps(s, want[:, 0]).max() <= 1 and _ulps(c, want[:, 1]).max() <= 1
