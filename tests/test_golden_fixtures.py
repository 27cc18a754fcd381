"""tests/golden/*.npy: rows committed in an earlier round must be reproduced bit for bit by the oracle
built today (CPU) and by the CUDA path (GPU). See tests/golden/README.md."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import gen_golden_fixtures as G  # noqa: E402

from bevy_firework_b200.workloads import SEED  # noqa: E402

NAMES = sorted(G.scenes())


def _fixture(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npy"))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_committed_rows(oracle, name):
    cols, spawners, frames = G.scenes()[name]
    w = oracle.OracleWorld(seed=SEED)
    rows = G.run(w, cols, spawners, frames)
    w.close()
    want = _fixture(name)
    assert len(rows) == len(want) > 500
    assert rows.tobytes() == want.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_path_reproduces_committed_rows(engine, name):
    cols, spawners, frames = G.scenes()[name]
    rows = G.run(engine, cols, spawners, frames)
    want = _fixture(name)
    assert len(rows) == len(want)
    for f in want.dtype.names:  # IEEE equality per field (a -0 / +0 of a per-stream constant may differ in sign)
        assert (rows[f] == want[f]).all(), (name, f)
