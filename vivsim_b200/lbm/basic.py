"""Import path of the reference's vivsim/lbm/basic.py: the same public names, implemented in vivsim_b200.lbm
(C ABI underneath, include/vivsim_b200.h)."""

from vivsim_b200.lbm import (  # noqa: F401
    streaming,
    get_macroscopic,
    get_equilibrium,
    collision_bgk,
    get_omega,
    get_velocity_correction,
)
