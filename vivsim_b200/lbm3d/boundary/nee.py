"""Import path of the reference's vivsim/lbm3d/boundary/nee.py: the same public names, implemented in vivsim_b200.lbm3d
(C ABI underneath, include/vivsim_b200.h)."""

from vivsim_b200.lbm3d import (  # noqa: F401
    boundary_nee,
    boundary_velocity_nee,
    boundary_pressure_nee,
    boundary_force_corrected_nee,
)
