"""vivsim_b200: B200-native (sm_100a) implementation of the IB-LBM time step of haimingz/vivsim.

``lbm`` / ``lbm3d`` / ``ib`` / ``ib3d`` / ``dyn`` / ``post`` / ``multigrid`` keep the reference's pure-function names;
``Stepper`` is the fused hot path.  Compute lives in libvivsim_b200.so (C ABI, include/vivsim_b200.h);
there is no CPU fallback."""

from . import dyn, ib, ib3d, lbm, lbm3d, multigrid, post  # noqa: F401
from ._lib import LIB_PATH, VsbError, lib  # noqa: F401
from .stepper import Ensemble, Stepper  # noqa: F401

__version__ = "0.1.0"
