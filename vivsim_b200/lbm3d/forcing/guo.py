"""Import path of the reference's vivsim/lbm3d/forcing/guo.py: the same public names, implemented in vivsim_b200.lbm3d
(C ABI underneath, include/vivsim_b200.h)."""

from vivsim_b200.lbm3d import (  # noqa: F401
    get_guo_forcing_term,
    forcing_guo_bgk,
    get_mrt_forcing_operator,
    forcing_guo_mrt,
)
