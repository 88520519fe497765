"""Import path of the reference's vivsim/ib3d/geometry.py: the same public names, implemented in vivsim_b200.ib3d
(C ABI underneath, include/vivsim_b200.h)."""

from vivsim_b200.ib3d import (  # noqa: F401
    get_triangle_areas,
    get_surface_area,
    get_volume,
    get_ds,
)
