"""The C++ host mirror (host/firework.hpp) drives the same C ABI: its sparks example must
reproduce the oracle's live counts frame by frame."""
import json
import os
import subprocess

import numpy as np
import pytest

from bevy_firework_b200._native import frame_input
from bevy_firework_b200.workloads import sparks_spawner

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DT = float(np.float32(1.0) / np.float32(60.0))


def test_cpp_sparks_example_matches_oracle(oracle):
    import __graft_entry__ as g

    g.build_cpp_host()
    out = subprocess.check_output([os.path.join(ROOT, "host", "bin", "sparks"), "150", "1000"], text=True, timeout=120)
    res = json.loads(out)
    w = oracle.OracleWorld(seed=0x00F12E00)
    ps, nt, es, ne = sparks_spawner(1000.0).pods()
    w.spawner_reset(res["entity"], ps, nt, es, ne, True)
    want = []
    for _ in range(150):
        w.frame(DT, [frame_input(res["entity"], (0.0, 0.1, 0.0))])
        want.append(w.counts(res["entity"])[0])
    assert res["counts"] == want
    rows = w.read_particles(res["entity"], 0)
    assert res["live"] == len(rows) and res["active"] == 1
    # sums over bit-equal rows in the same order, accumulated in double on both sides (printed with 9 digits)
    assert res["sum_age"] == pytest.approx(float(rows["age"].astype(np.float64).sum()), rel=2e-9)
    assert res["sum_y"] == pytest.approx(float(rows["position"][:, 1].astype(np.float64).sum()), rel=2e-9)


def test_c99_sparks_example_matches_oracle(oracle):
    """the same scene through the bare C ABI from a C99 program (host/examples/sparks_c99.c)"""
    import __graft_entry__ as g

    g.build_cpp_host()
    out = subprocess.check_output([os.path.join(ROOT, "host", "bin", "sparks_c99"), "150", "1000"], text=True, timeout=120)
    res = json.loads(out)
    w = oracle.OracleWorld(seed=0x00F12E00)
    ps, nt, es, ne = sparks_spawner(1000.0).pods()
    w.spawner_reset(1, ps, nt, es, ne, True)
    want = []
    for _ in range(150):
        w.frame(DT, [frame_input(1, (0.0, 0.1, 0.0))])
        want.append(w.counts(1)[0])
    assert res["counts"] == want and res["bytes_per_particle"] == 80
    rows = w.read_particles(1, 0)
    assert res["live"] == len(rows)
    assert res["sum_age"] == pytest.approx(float(rows["age"].astype(np.float64).sum()), rel=2e-9)
    assert res["sum_y"] == pytest.approx(float(rows["position"][:, 1].astype(np.float64).sum()), rel=2e-9)


def test_cpp_stress_example_runs():
    out = subprocess.check_output([os.path.join(ROOT, "host", "bin", "stress_test"), "64", "15625", "50"], text=True, timeout=120)
    res = json.loads(out)
    assert 950_000 < res["live_particles"] < 1_010_000 and res["particles_per_s"] > 1e9
