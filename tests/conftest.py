import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Native artefacts (CUDA library, oracle) compiled in-tree."""
    import __graft_entry__ as g

    g.build_cuda()
    import oracle

    oracle.build()
    return True


@pytest.fixture(scope="session")
def params06():
    from quadruped_control_b200 import default_params

    return default_params(0.6)


@pytest.fixture(scope="session")
def params08():
    from quadruped_control_b200 import default_params

    return default_params(0.8)


@pytest.fixture(scope="session")
def solver06(built, params06):
    from quadruped_control_b200 import lib

    s = lib.BalanceSolver(params06, device=0)
    yield s
    s.close()


def rel_err(a, b, floor=1.0):
    """max over batch of ||a-b||_inf / max(||b||_inf, floor)  (SURVEY.md 8d metric)."""
    import numpy as np

    a = np.asarray(a).reshape(len(a), -1)
    b = np.asarray(b).reshape(len(b), -1)
    if a.shape[0] == 0:
        return 0.0
    return float((np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), floor)).max())
