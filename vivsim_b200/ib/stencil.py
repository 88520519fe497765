"""Import path of the reference's vivsim/ib/stencil.py: the same public names, implemented in vivsim_b200.ib
(C ABI underneath, include/vivsim_b200.h)."""

from vivsim_b200.ib import (  # noqa: F401
    get_ib_stencil,
    interpolate,
    spread,
)
