"""Import path of the reference's vivsim/lbm/forcing/edm.py: the same public names, implemented in vivsim_b200.lbm
(C ABI underneath, include/vivsim_b200.h)."""

from vivsim_b200.lbm import (  # noqa: F401
    forcing_edm,
)
