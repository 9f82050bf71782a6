import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
        return cache[name]
    return load


@pytest.fixture(autouse=True)
def _restore_math_mode():
    """Tests that switch the math mode (set_math) leave the package default behind them."""
    try:
        import neuralnetlib_b200 as nn
    except Exception:   # noqa: BLE001
        yield
        return
    prev = nn.get_math()
    yield
    nn.set_math(prev)
