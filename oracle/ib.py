"""Immersed-boundary oracle with the reference's ``vivsim.ib`` names
(NumPy fp32; test infrastructure only)."""

import numpy as np

from .core import F32, f32


# ------------------------------------------------------------------ a17 delta kernels
def kernel_peskin_3pt(r):
    """Support 1.5.  Reference: ib/kernels.py:4-22."""
    a = np.abs(f32(r))
    with np.errstate(invalid="ignore"):
        inner = (F32(1) + np.sqrt(F32(1) - F32(3) * a * a)) / F32(3)
        outer = (F32(5) - F32(3) * a - np.sqrt(F32(-2) + F32(6) * a - F32(3) * a * a)) / F32(6)
    return np.where(a > 1.5, F32(0), np.where(a < 0.5, inner, outer)).astype(F32)


def kernel_peskin_4pt(r):
    """Support 2.  Reference: ib/kernels.py:25-43."""
    a = np.abs(f32(r))
    with np.errstate(invalid="ignore"):
        inner = (F32(3) - F32(2) * a + np.sqrt(F32(1) + F32(4) * a - F32(4) * a * a)) * F32(0.125)
        outer = (F32(5) - F32(2) * a - np.sqrt(F32(-7) + F32(12) * a - F32(4) * a * a)) * F32(0.125)
    return np.where(a > 2, F32(0), np.where(a < 1, inner, outer)).astype(F32)


def kernel_cosine_4pt(r):
    """(1 + cos(pi r / 2)) / 4 on support 2.  Reference: ib/kernels.py:46-61."""
    a = np.abs(f32(r))
    return np.where(a > 2, F32(0), (F32(1) + np.cos(F32(np.pi) * a * F32(0.5))) * F32(0.25)).astype(F32)


def kernel_hat_2pt(r):
    """Standard 2-point hat max(0, 1-|r|).  NOT in the reference (README lists a
    2-point kernel, ib/kernels.py does not define one) -- unpinned extra."""
    return np.maximum(F32(0), F32(1) - np.abs(f32(r))).astype(F32)


# ------------------------------------------------------------------ a18 stencil
def _offsets(radius):
    return np.arange(-radius + 1, radius + 1, dtype=np.int32)


def get_ib_stencil(marker_x, marker_y, ny, kernel=kernel_peskin_4pt, stencil_radius=2):
    """Tensor-product weights and flat indices ``x*ny + y`` (no wrap / clamp).
    Reference: ib/stencil.py:27-51."""
    mx = f32(marker_x); my = f32(marker_y)
    off = _offsets(stencil_radius)
    ox, oy = (o.reshape(1, -1) for o in np.meshgrid(off, off, indexing="ij"))
    sx = np.floor(mx).astype(np.int32)[:, None] + ox
    sy = np.floor(my).astype(np.int32)[:, None] + oy
    w = kernel(sx.astype(F32) - mx[:, None]) * kernel(sy.astype(F32) - my[:, None])
    return w.astype(F32), (sx * np.int32(ny) + sy).astype(np.int32)


# ------------------------------------------------------------------ a19 / a20
def interpolate(grid_values, stencil_weights, stencil_indices):
    """out[m,c] = sum_s w[m,s] grid[c, idx[m,s]].  Reference: ib/stencil.py:76-78."""
    g = f32(grid_values)
    flat = g.reshape(g.shape[0], -1)
    vals = flat[:, stencil_indices]                    # (C, M, S)
    return np.einsum("ms,cms->mc", f32(stencil_weights), vals).astype(F32)


def spread(marker_values, grid_values, stencil_weights, stencil_indices):
    """grid[c, idx[m,s]] += val[m,c] w[m,s]; duplicates accumulate.
    Reference: ib/stencil.py:104-110."""
    g = f32(grid_values)
    flat = g.reshape(g.shape[0], -1).copy()
    contrib = np.einsum("mc,ms->cms", f32(marker_values), f32(stencil_weights)).astype(F32)
    np.add.at(flat, (slice(None), stencil_indices), contrib)
    return flat.reshape(g.shape)


# ------------------------------------------------------------------ a21 MDF
def multi_direct_forcing(grid_u, stencil_weights, stencil_indices, marker_u_target, marker_ds, n_iter=5):
    """Multi-direct forcing.  Reference: ib/mdf.py:31-64 (rho = 1 assumed in the 0.5)."""
    grid_u = f32(grid_u)
    target = f32(marker_u_target)
    ds2 = f32(marker_ds).reshape(-1, 1) * F32(2)
    zero = np.zeros_like(grid_u)
    total = np.zeros_like(target)
    um = interpolate(grid_u, stencil_weights, stencil_indices)
    for _ in range(n_iter):
        step = (target - um) * ds2
        total = total + step
        um = um + interpolate(spread(step, zero, stencil_weights, stencil_indices) * F32(0.5),
                              stencil_weights, stencil_indices)
    return spread(total, zero, stencil_weights, stencil_indices), -total


# ------------------------------------------------------------------ geometry (fixtures)
def get_area(marker_coords):
    """Shoelace area.  Reference: ib/geometry.py:15-18."""
    p = f32(marker_coords)
    x, y = p[:, 0], p[:, 1]
    return F32(0.5) * np.abs(np.sum(x * np.roll(y, 1) - y * np.roll(x, 1), dtype=F32))


def get_ds(marker_coords, closed=True):
    """Arc-length weight per marker.  Reference: ib/geometry.py:31-43."""
    p = f32(marker_coords)
    if closed:
        seg = np.linalg.norm(p - np.roll(p, -1, axis=0), axis=1).astype(F32)
        return ((seg + np.roll(seg, 1)) / F32(2)).astype(F32)
    seg = np.linalg.norm(p[1:] - p[:-1], axis=1).astype(F32) / F32(2)
    return (np.pad(seg, (1, 0)) + np.pad(seg, (0, 1))).astype(F32)
