import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    """One device context shared by the GPU tests (created lazily; fails loudly without a GPU)."""
    import pmp_mcmc_b200 as pm
    c = pm.Context(device=0)
    yield c
    c.close()


def synthetic_linear(n, seed=0):
    """x ~ U(-1,1), y = -1 + 2x + 0.5 eps (lb.py:11-17, convery_time_MP.cu:107-110), numpy Generator(seed)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, n).astype(np.float32)
    y = (-1.0 + 2.0 * x + 0.5 * rng.standard_normal(n)).astype(np.float32)
    return x, y
