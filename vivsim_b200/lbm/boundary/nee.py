"""Import path of the reference's vivsim/lbm/boundary/nee.py: the same public names, implemented in vivsim_b200.lbm
(C ABI underneath, include/vivsim_b200.h)."""

from vivsim_b200.lbm import (  # noqa: F401
    boundary_nee,
    boundary_velocity_nee,
    boundary_pressure_nee,
    boundary_force_corrected_nee,
)
