import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def ctx():
    """The library context on cuda:0 (GPU tests only)."""
    from mdcraft_b200 import _capi

    return _capi.Context(0)


def universe_from(pos, dims, **kw):
    from mdcraft_b200 import synthetic

    traj = synthetic.MemoryTrajectory(np.asarray(pos, np.float32), None if dims is None else np.asarray(dims, np.float32))
    return synthetic.Universe(traj, **kw)
