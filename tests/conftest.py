import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# north_star tolerance: fp32 results agree within 1e-5 relative (to the field's
# scale) per step and over a 100-step horizon; permutations are bit-exact.
RTOL = 1e-5


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def assert_close(actual, expected, rtol=RTOL, what=""):
    a = np.asarray(actual, dtype=np.float64)
    e = np.asarray(expected, dtype=np.float64)
    assert a.shape == e.shape, f"{what}: shape {a.shape} != {e.shape}"
    assert np.isfinite(a).all(), f"{what}: non-finite values"
    scale = max(float(np.abs(e).max()), 1e-30)
    err = float(np.abs(a - e).max()) / scale
    assert err <= rtol, f"{what}: max |diff| / max |ref| = {err:.3e} > {rtol:g}"


def assert_bitexact(actual, expected, what=""):
    a = np.asarray(actual)
    e = np.asarray(expected)
    assert a.shape == e.shape and a.dtype == e.dtype, f"{what}: {a.shape}/{a.dtype} vs {e.shape}/{e.dtype}"
    assert np.array_equal(a.view(np.uint8), e.view(np.uint8)), f"{what}: not bit-identical"


@pytest.fixture(scope="session")
def golden():
    return {n: load_golden(n) for n in ("lattice", "ops2d", "ops3d", "ib", "dyn", "recipes")}
