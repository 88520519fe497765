"""Import path of the reference's vivsim/lbm/boundary/eq.py: the same public names, implemented in vivsim_b200.lbm
(C ABI underneath, include/vivsim_b200.h)."""

from vivsim_b200.lbm import (  # noqa: F401
    boundary_equilibrium,
    boundary_velocity_equilibrium,
    boundary_pressure_equilibrium,
    boundary_force_corrected_equilibrium,
)
