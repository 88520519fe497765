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


def rel_err(actual, expected):
    a = np.asarray(actual, dtype=np.float64)
    e = np.asarray(expected, dtype=np.float64)
    return float(np.abs(a - e).max()) / max(float(np.abs(e).max()), 1e-30)


def assert_within_fp32_drift(actual, oracle32, truth64, what="", factor=3.0, floor=RTOL):
    """Long-horizon check against an fp64 yardstick.  Two correct fp32 evaluations of the same recipe (different
    summation orders: atomics, BLAS, plain loops) drift apart over many steps, most visibly in the marker forces
    (U - u_m) 2 ds, a difference of nearly equal numbers.  `truth64` is the same algorithm in double precision
    (oracle.cport with dtype=float64), `oracle32` the fp32 oracle.  The CUDA result passes when its distance from the
    fp64 result is within `factor` x the fp32 oracle's own distance from it -- i.e. it is as good an fp32 evaluation
    as the oracle is -- or within the 1e-5 of north_star, whichever is larger."""
    a = np.asarray(actual, dtype=np.float64)
    assert np.isfinite(a).all(), f"{what}: non-finite values"
    drift = rel_err(oracle32, truth64)
    err = rel_err(a, truth64)
    bound = max(floor, factor * drift)
    assert err <= bound, (f"{what}: |cuda - fp64| = {err:.3e} exceeds max({floor:g}, {factor:g} x |fp32 oracle - fp64| "
                          f"= {factor * drift:.3e})")
    return err, drift


def assert_bitexact(actual, expected, what=""):
    a = np.asarray(actual)
    e = np.asarray(expected)
    assert a.shape == e.shape and a.dtype == e.dtype, f"{what}: {a.shape}/{a.dtype} vs {e.shape}/{e.dtype}"
    assert np.array_equal(a.view(np.uint8), e.view(np.uint8)), f"{what}: not bit-identical"


@pytest.fixture(scope="session")
def golden():
    return {n: load_golden(n) for n in ("lattice", "ops2d", "ops3d", "ib", "dyn", "recipes", "rotation")}
