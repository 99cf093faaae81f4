import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "data")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle backend (test infrastructure; see oracle/)."""
    from oracle import orc
    return orc


@pytest.fixture(scope="session")
def orc_backend(oracle):
    return oracle.backend()


@pytest.fixture(scope="session")
def gpu_backend():
    from fountain_b200.api import default_backend
    return default_backend()


@pytest.fixture(scope="session")
def rounded_cube_path():
    return os.path.join(DATA, "rounded_cube.ply")


def unit_sphere_dirs(n, seed):
    """Uniform directions (UnitSphereSurface in tests/tri_watertight.rs:27), seeded."""
    rng = np.random.default_rng(seed)
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v.astype(np.float32)
