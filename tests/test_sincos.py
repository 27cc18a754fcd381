"""include/fw_sincos.h -- the sine / cosine both sides of every parity comparison compile.

It replaces a platform libm call of the reference (glam -> f32::sin_cos), so it is pinned on its
own: against a 200-bit evaluation (mpmath) it must be the correctly rounded float, or within
0.5 + 2^-12 ulp (the Taylor truncation bound stated in the header), on every class of argument --
the fast double reduction, the Payne-Hanek path for |x| >= 2^20, arguments next to multiples of
pi/2, denormals, zeros, non-finite values. scripts/sincos_exhaustive.c runs all 2^32 bit patterns
against the 80-bit x87 libm (result committed: profiles/r2/sincos_exhaustive.txt)."""
import math

import mpmath as mp
import numpy as np
import pytest

f32 = np.float32


def _ulp_err(got: float, want) -> float:
    if want == 0:
        return 0.0 if got == 0 else float("inf")
    e = int(mp.floor(mp.log(abs(want), 2)))
    ulp = mp.mpf(2) ** max(e - 23, -149)
    return float(abs(mp.mpf(float(got)) - want) / ulp)


def _samples():
    rng = np.random.default_rng(20261017)
    xs = [rng.uniform(-8, 8, 4000), rng.uniform(-1e3, 1e3, 2000), rng.uniform(-1e6, 1e6, 2000),
          rng.uniform(-2e6, 2e6, 500), rng.uniform(-1e-3, 1e-3, 500)]
    bits = rng.integers(0, 2 ** 32, 6000, dtype=np.uint64).astype(np.uint32)  # any float, incl. huge / denormal
    xs.append(bits.view(np.float32).astype(np.float64))
    k = rng.integers(-700000, 700000, 3000)
    near = (k * (math.pi / 2)).astype(np.float32)  # floats next to multiples of pi/2: worst cancellation
    xs.append(near.astype(np.float64))
    xs.append(np.nextafter(near, f32(np.inf)).astype(np.float64))
    xs.append(np.array([2.0 ** 20, np.nextafter(f32(2.0 ** 20), f32(0)), 2.0 ** 24, 1e22, 3.4028234e38, 1e-45, 1.17549435e-38,
                        float.fromhex('0x1.ca793ap+22'), float.fromhex('0x1.2776fep+33'), math.pi, math.pi / 2, math.pi / 4]))
    x = np.concatenate(xs).astype(np.float32)
    x = x[np.isfinite(x)]
    return np.concatenate([x, -x[:2000]])


def test_correctly_rounded_against_200_bits(oracle):
    mp.mp.prec = 220
    x = _samples()
    s, c = oracle.sincosf(x)
    worst = 0.0
    misrounded = 0
    for xi, si, ci in zip(x.tolist(), s.tolist(), c.tolist()):
        ws, wc = mp.sin(mp.mpf(xi)), mp.cos(mp.mpf(xi))
        for got, want in ((si, ws), (ci, wc)):
            err = _ulp_err(got, want)
            worst = max(worst, err)
            misrounded += got != float(f32(float(want)))
    assert worst <= 0.5 + 2.0 ** -12, worst
    assert misrounded <= len(x) // 1000, misrounded  # (exhaustive run: 52 of 8.6e9 results)


def test_special_values(oracle):
    s, c = oracle.sincosf(0.0)
    assert s == 0.0 and math.copysign(1.0, s) == 1.0 and c == 1.0
    s, c = oracle.sincosf(-0.0)
    assert s == 0.0 and math.copysign(1.0, s) == -1.0 and c == 1.0  # sin(-0) = -0
    for bad in (float("inf"), float("-inf"), float("nan")):
        s, c = oracle.sincosf(bad)
        assert math.isnan(s) and math.isnan(c)
    # odd / even symmetry holds bit for bit
    x = np.random.default_rng(3).uniform(-1e7, 1e7, 20000).astype(np.float32)
    s, c = oracle.sincosf(x)
    s2, c2 = oracle.sincosf(-x)
    assert (s2 == -s).all() and (c2 == c).all()
    # sin^2 + cos^2 = 1 to float rounding, everywhere
    assert np.abs(s.astype(np.float64) ** 2 + c.astype(np.float64) ** 2 - 1.0).max() < 2.5e-7


def test_exhaustive_run_is_committed():
    import os

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2", "sincos_exhaustive.txt")
    text = open(path).read()
    import re

    assert "floats tested 4278190080" in text, text  # every finite float
    worst = [float(v) for v in re.findall(r"worst ([0-9.]+) ulp", text)]
    wrong = [int(v) for v in re.findall(r"not correctly rounded (\d+)", text)]
    assert len(worst) == 2 and max(worst) <= 0.5 + 2.0 ** -12 and max(wrong) <= 64, text


@pytest.mark.gpu
def test_device_equals_host_bit_for_bit(engine, oracle):
    """the same header compiled by nvcc for sm_100a (-fmad=false) and by gcc (-ffp-contract=off):
    every result bit equal, on both reduction paths"""
    x = _samples()
    x = np.concatenate([x, np.array([np.inf, -np.inf, np.nan], dtype=np.float32)])
    hs, hc = oracle.sincosf(x)
    ds, dc = engine.device_sincos(x)
    assert (hs.view(np.uint32) == ds.view(np.uint32))[np.isfinite(hs)].all()
    assert (hc.view(np.uint32) == dc.view(np.uint32))[np.isfinite(hc)].all()
    assert (np.isnan(hs) == np.isnan(ds)).all() and (np.isnan(hc) == np.isnan(dc)).all()
