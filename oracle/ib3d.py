"""3-D immersed-boundary oracle (``vivsim.ib3d`` names; test infrastructure only).

interpolate / spread / multi_direct_forcing / kernels are dimension-agnostic and
re-exported from the 2-D module exactly as the reference does
(ib3d/__init__.py:10-12)."""

import numpy as np

from .core import F32, f32
from .ib import (kernel_cosine_4pt, kernel_hat_2pt, kernel_peskin_3pt, kernel_peskin_4pt,  # noqa: F401
                 interpolate, spread, multi_direct_forcing, _offsets)


def get_ib_stencil(marker_coords, grid_shape, kernel=kernel_peskin_4pt, stencil_radius=2):
    """(2r)^3 tensor-product stencil, flat index x*ny*nz + y*nz + z.
    Reference: ib3d/stencil.py:26-59 (ValueError on bad shapes, :26-34)."""
    if len(grid_shape) != 3:
        raise ValueError(f"grid_shape must be a 3-tuple, got {grid_shape}.")
    p = f32(marker_coords)
    if p.ndim != 2 or p.shape[1] != 3:
        raise ValueError(f"marker_coords must have shape (n_markers, 3), got {p.shape}.")
    _, ny, nz = grid_shape
    off = _offsets(stencil_radius)
    ox, oy, oz = (o.reshape(1, -1) for o in np.meshgrid(off, off, off, indexing="ij"))
    base = np.floor(p).astype(np.int32)
    sx, sy, sz = base[:, 0:1] + ox, base[:, 1:2] + oy, base[:, 2:3] + oz
    w = (kernel(sx.astype(F32) - p[:, 0:1]) * kernel(sy.astype(F32) - p[:, 1:2])
         * kernel(sz.astype(F32) - p[:, 2:3]))
    return w.astype(F32), (sx * np.int32(ny * nz) + sy * np.int32(nz) + sz).astype(np.int32)


def get_triangle_areas(vertex_coords, faces):
    """Reference: ib3d/geometry.py:16-19."""
    t = f32(vertex_coords)[np.asarray(faces)]
    return (F32(0.5) * np.linalg.norm(np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0]), axis=1)).astype(F32)


def get_surface_area(vertex_coords, faces):
    """Reference: ib3d/geometry.py:24."""
    return np.sum(get_triangle_areas(vertex_coords, faces), dtype=F32)


def get_volume(vertex_coords, faces):
    """Reference: ib3d/geometry.py:40-46."""
    t = f32(vertex_coords)[np.asarray(faces)]
    vol = np.einsum("ij,ij->i", t[:, 0], np.cross(t[:, 1], t[:, 2])).astype(F32) / F32(6)
    return np.abs(np.sum(vol, dtype=F32))


def get_ds(vertex_coords, faces):
    """Lumped vertex area (a third of each incident triangle).  Reference: ib3d/geometry.py:61-67."""
    v = f32(vertex_coords)
    faces = np.asarray(faces)
    out = np.zeros(v.shape[0], dtype=F32)
    np.add.at(out, faces.reshape(-1), np.repeat(get_triangle_areas(v, faces) / F32(3), 3))
    return out


def icosphere(radius, center, subdivisions):
    """Icosphere fixture generator (same construction as examples/benchmark3d.py:30-105,
    examples/3d/flow_past_sphere.py:44-89): 10*4^n + 2 vertices."""
    phi = (1 + 5 ** 0.5) / 2
    v = [(-1, phi, 0), (1, phi, 0), (-1, -phi, 0), (1, -phi, 0), (0, -1, phi), (0, 1, phi),
         (0, -1, -phi), (0, 1, -phi), (phi, 0, -1), (phi, 0, 1), (-phi, 0, -1), (-phi, 0, 1)]
    verts = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    tris = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4),
            (11, 10, 2), (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8),
            (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    for _ in range(subdivisions):
        cache, nxt = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]

        for a, b, c in tris:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nxt += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        tris = nxt
    pts = np.array(verts) * radius + np.asarray(center, dtype=np.float64)
    return pts.astype(F32), np.array(tris, dtype=np.int32)
