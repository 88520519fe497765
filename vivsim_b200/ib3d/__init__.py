"""3-D immersed-boundary operators with the names of the reference's ``vivsim.ib3d``
(vivsim/ib3d/__init__.py:3-12).  interpolate / spread / multi_direct_forcing and the delta
kernels are dimension-agnostic and shared with ``vivsim_b200.ib``, as in the reference."""

import numpy as np
import torch

from .. import _lib as L
from ..ib import (kernel_cosine_4pt, kernel_hat_2pt, kernel_peskin_3pt, kernel_peskin_4pt,  # noqa: F401
                  interpolate, spread, multi_direct_forcing, _stencil, _np32)


def get_ib_stencil(marker_coords, grid_shape, kernel=kernel_peskin_4pt, stencil_radius=2):
    """(n_markers, (2r)^3) weights and flat indices x*ny*nz + y*nz + z   (ib3d/stencil.py:7-59)."""
    if len(grid_shape) != 3:
        raise ValueError(f"grid_shape must be a 3-tuple, got {grid_shape}.")
    if not hasattr(marker_coords, "ndim") or marker_coords.ndim != 2 or marker_coords.shape[1] != 3:
        raise ValueError(f"marker_coords must have shape (n_markers, 3), got {tuple(getattr(marker_coords, 'shape', ()))}.")
    _, ny, nz = grid_shape
    return _stencil(3, marker_coords, ny, nz, kernel, stencil_radius)


# ---- setup-time geometry on the host (ib3d/geometry.py) ----
def get_triangle_areas(vertex_coords, faces):
    """ib3d/geometry.py:6-19."""
    t = _np32(vertex_coords)[_faces(faces)]
    return (np.float32(0.5) * np.linalg.norm(np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0]), axis=1)).astype(np.float32)


def get_surface_area(vertex_coords, faces):
    """ib3d/geometry.py:22-24."""
    return np.sum(get_triangle_areas(vertex_coords, faces), dtype=np.float32)


def get_volume(vertex_coords, faces):
    """ib3d/geometry.py:27-46."""
    t = _np32(vertex_coords)[_faces(faces)]
    signed = np.einsum("ij,ij->i", t[:, 0], np.cross(t[:, 1], t[:, 2])).astype(np.float32) / np.float32(6)
    return np.abs(np.sum(signed, dtype=np.float32))


def get_ds(vertex_coords, faces):
    """Lumped vertex areas   (ib3d/geometry.py:49-67)."""
    v = _np32(vertex_coords)
    fc = _faces(faces)
    out = np.zeros(v.shape[0], dtype=np.float32)
    np.add.at(out, fc.reshape(-1), np.repeat(get_triangle_areas(v, fc) / np.float32(3), 3))
    return out


def _faces(faces):
    if isinstance(faces, torch.Tensor):
        faces = faces.detach().cpu().numpy()
    return np.asarray(faces)
