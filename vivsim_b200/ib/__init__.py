"""Immersed-boundary operators with the names of the reference's ``vivsim.ib``
(vivsim/ib/__init__.py:3-6) on CUDA tensors."""

import ctypes as C

import numpy as np
import torch

from .. import _lib as L


def _delta(kind):
    def kernel(r):
        r = L.dev(r, name="r")
        out = torch.empty_like(r)
        L.check(L.lib().vsb_ib_delta(L.DELTA[kind], C.c_int64(r.numel()), L.ptr(r), L.ptr(out), L.stream()))
        return out
    kernel.delta_kind = kind
    return kernel


kernel_peskin_3pt = _delta("peskin3")   # ib/kernels.py:4-22
kernel_peskin_4pt = _delta("peskin4")   # ib/kernels.py:25-43
kernel_cosine_4pt = _delta("cosine4")   # ib/kernels.py:46-61
kernel_hat_2pt = _delta("hat2")         # named in the reference README, not defined there (unpinned extra)


def delta_kind(kernel):
    """Map a kernel callable (or name) to the C enum; the reference takes a Python callable
    (ib/stencil.py:11), which cannot cross the ABI, so unknown callables are rejected."""
    if isinstance(kernel, str) and kernel in L.DELTA:
        return kernel
    kind = getattr(kernel, "delta_kind", None)
    if kind not in L.DELTA:
        raise ValueError("kernel must be one of kernel_peskin_3pt, kernel_peskin_4pt, kernel_cosine_4pt, kernel_hat_2pt")
    return kind


def _stencil(dim, coords, ny, nz, kernel, stencil_radius):
    coords = L.dev(coords, name="marker coordinates")
    m = coords.shape[0]
    ns = (2 * stencil_radius) ** dim
    w = torch.empty((m, ns), device=coords.device, dtype=torch.float32)
    idx = torch.empty((m, ns), device=coords.device, dtype=torch.int32)
    L.check(L.lib().vsb_ib_stencil(dim, L.DELTA[delta_kind(kernel)], int(stencil_radius), C.c_int64(m), L.ptr(coords),
                                   int(ny), int(nz), L.ptr(w), L.ptr(idx), L.stream()))
    return w, idx


def get_ib_stencil(marker_x, marker_y, ny, kernel=kernel_peskin_4pt, stencil_radius=2):
    """Stencil weights and flat indices x*ny + y, both (n_markers, (2r)^2)   (ib/stencil.py:7-51)."""
    mx = L.dev(marker_x, name="marker_x")
    my = L.dev(marker_y, name="marker_y")
    if mx.ndim != 1 or mx.shape != my.shape:
        raise ValueError("marker_x and marker_y must be 1-D arrays of the same length")
    return _stencil(2, torch.stack([mx, my], dim=1), ny, 1, kernel, stencil_radius)


def interpolate(grid_values, stencil_weights, stencil_indices):
    """(C, *grid) -> (n_markers, C)   (ib/stencil.py:54-78)."""
    g = L.dev(grid_values, name="grid_values")
    w = L.dev(stencil_weights, name="stencil_weights")
    idx = L.dev(stencil_indices, torch.int32, name="stencil_indices")
    if w.shape != idx.shape or w.ndim != 2:
        raise ValueError("stencil_weights and stencil_indices must share a (n_markers, n_stencil) shape")
    out = torch.empty((w.shape[0], g.shape[0]), device=g.device, dtype=torch.float32)
    L.check(L.lib().vsb_ib_interpolate(int(g.shape[0]), C.c_int64(g[0].numel()), L.ptr(g), C.c_int64(w.shape[0]),
                                       int(w.shape[1]), L.ptr(w), L.ptr(idx), L.ptr(out), L.stream()))
    return out


def spread(marker_values, grid_values, stencil_weights, stencil_indices, ordered=False):
    """Scatter-add marker values onto a copy of grid_values   (ib/stencil.py:81-110).
    ordered=True: deterministic summation order (flattened (marker, stencil point) order per cell, as a sequential
    scatter applies it): reproducible run to run and bit-identical to the CPU oracle; ordered=False: fp32 atomics."""
    v = L.dev(marker_values, name="marker_values")
    g = L.dev(grid_values, name="grid_values").clone()
    w = L.dev(stencil_weights, name="stencil_weights")
    idx = L.dev(stencil_indices, torch.int32, name="stencil_indices")
    if v.shape != (w.shape[0], g.shape[0]):
        raise ValueError(f"marker_values must have shape ({w.shape[0]}, {g.shape[0]}), got {tuple(v.shape)}")
    if ordered:
        lib = L.lib()
        lib.vsb_ib_spread_ordered_workspace.restype = C.c_int64
        need = int(lib.vsb_ib_spread_ordered_workspace(C.c_int64(w.shape[0]), int(w.shape[1])))
        if need < 0:
            raise ValueError("ordered spread: more than 2^31 stencil entries")
        ws = torch.empty(max(need, 1), dtype=torch.uint8, device=g.device)
        L.check(lib.vsb_ib_spread_ordered(int(g.shape[0]), C.c_int64(g[0].numel()), L.ptr(g), C.c_int64(w.shape[0]),
                                          int(w.shape[1]), L.ptr(v), L.ptr(w), L.ptr(idx), L.ptr(ws), C.c_int64(need),
                                          L.stream()))
        return g
    L.check(L.lib().vsb_ib_spread(int(g.shape[0]), C.c_int64(g[0].numel()), L.ptr(g), C.c_int64(w.shape[0]),
                                  int(w.shape[1]), L.ptr(v), L.ptr(w), L.ptr(idx), L.stream()))
    return g


def multi_direct_forcing(grid_u, stencil_weights, stencil_indices, marker_u_target, marker_ds, n_iter=5):
    """Multi-direct forcing with a precomputed stencil   (ib/mdf.py:10-64).
    Returns (grid_force, marker_reaction_force).  The fused stepper uses the on-the-fly kernel
    (vsb_ib_mdf) instead; this form exists for drop-in parity with the reference signature."""
    grid_u = L.dev(grid_u, name="grid_u")
    target = L.dev(marker_u_target, name="marker_u_target")
    ds = marker_ds if isinstance(marker_ds, torch.Tensor) else torch.as_tensor(marker_ds, device=grid_u.device)
    ds2 = L.dev(ds, name="marker_ds").reshape(-1, 1) * 2.0
    zero = torch.zeros_like(grid_u)
    total = torch.zeros_like(target)
    um = interpolate(grid_u, stencil_weights, stencil_indices)
    for _ in range(int(n_iter)):
        step = (target - um) * ds2
        total = total + step
        um = um + interpolate(spread(step, zero, stencil_weights, stencil_indices) * 0.5, stencil_weights, stencil_indices)
    return spread(total, zero, stencil_weights, stencil_indices), -total


# ---- setup-time geometry: O(markers), host NumPy (out of scope as kernels) ----
def get_area(marker_coords):
    """Shoelace area of a closed polygon   (ib/geometry.py:6-18)."""
    p = _np32(marker_coords)
    x, y = p[:, 0], p[:, 1]
    return np.float32(0.5) * np.abs(np.sum(x * np.roll(y, 1) - y * np.roll(x, 1), dtype=np.float32))


def get_ds(marker_coords, closed=True):
    """Arc-length weight of each marker   (ib/geometry.py:21-43)."""
    p = _np32(marker_coords)
    if closed:
        seg = np.linalg.norm(p - np.roll(p, -1, axis=0), axis=1).astype(np.float32)
        return ((seg + np.roll(seg, 1)) / np.float32(2)).astype(np.float32)
    seg = np.linalg.norm(p[1:] - p[:-1], axis=1).astype(np.float32) / np.float32(2)
    return (np.pad(seg, (1, 0)) + np.pad(seg, (0, 1))).astype(np.float32)


def _np32(x):
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.asarray(x, dtype=np.float32)
