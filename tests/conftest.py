import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.lib()
    return O


@pytest.fixture()
def engine():
    """A fresh fw_context on cuda:0 through the C ABI (fails loudly without the library/GPU)."""
    from bevy_firework_b200._native import Engine

    eng = Engine(device=0, seed=0x00F12E00)
    yield eng
    eng.close()
