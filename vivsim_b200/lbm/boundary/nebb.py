"""Import path of the reference's vivsim/lbm/boundary/nebb.py: the same public names, implemented in vivsim_b200.lbm
(C ABI underneath, include/vivsim_b200.h)."""

from vivsim_b200.lbm import (  # noqa: F401
    boundary_nebb,
    boundary_velocity_nebb,
    boundary_pressure_nebb,
    boundary_force_corrected_nebb,
)
