"""Import path of the reference's vivsim/lbm3d/forcing/edm.py: the same public names, implemented in vivsim_b200.lbm3d
(C ABI underneath, include/vivsim_b200.h)."""

from vivsim_b200.lbm3d import (  # noqa: F401
    forcing_edm,
)
