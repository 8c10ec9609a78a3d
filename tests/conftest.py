import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def gold_ll():
    return np.load(os.path.join(GOLDEN, "loglike_golden.npz"))


@pytest.fixture(scope="session")
def gold_replay():
    return np.load(os.path.join(GOLDEN, "replay_golden.npz"))


@pytest.fixture(scope="session")
def gold_pull():
    return np.load(os.path.join(GOLDEN, "pull_golden.npz"))


@pytest.fixture(scope="session")
def gold_radial():
    return np.load(os.path.join(GOLDEN, "radial_golden.npz"))


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import __graft_entry__ as g
    g.build()
    from mcdiff_b200.engine import Engine
    return Engine(0)
