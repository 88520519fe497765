"""Import path of the reference's vivsim/ib/kernels.py: the same public names, implemented in vivsim_b200.ib
(C ABI underneath, include/vivsim_b200.h)."""

from vivsim_b200.ib import (  # noqa: F401
    kernel_peskin_3pt,
    kernel_peskin_4pt,
    kernel_cosine_4pt,
)
