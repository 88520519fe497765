"""Import path of the reference's vivsim/ib3d/stencil.py: the same public names, implemented in vivsim_b200.ib3d
(C ABI underneath, include/vivsim_b200.h)."""

from vivsim_b200.ib3d import (  # noqa: F401
    get_ib_stencil,
)
