"""Rigid-body dynamics oracle (``vivsim.dyn`` names; NumPy fp32; test infrastructure only)."""

import numpy as np

from .core import F32, f32


def newmark(a, v, d, h, m, k, c, dt=1, gamma=0.5, beta=0.25):
    """One Newmark-beta step for scalar or matrix (m, k, c).  Reference: dyn.py:27-51."""
    a, v, d, h = f32(a), f32(v), f32(d), f32(h)
    c1, c2 = gamma * dt, beta * dt ** 2
    v_pred = v + F32(dt * (1 - gamma)) * a
    d_pred = d + F32(dt) * v + F32(dt ** 2 * (0.5 - beta)) * a
    if np.ndim(m) > 0:
        m, k, c = f32(m), f32(k), f32(c)
        a_new = np.linalg.solve(m + F32(c1) * c + F32(c2) * k, h - c @ v_pred - k @ d_pred).astype(F32)
    else:
        a_new = ((h - F32(c) * v_pred - F32(k) * d_pred) / F32(m + c1 * c + c2 * k)).astype(F32)
    return a_new, (F32(c1) * a_new + v_pred).astype(F32), (F32(c2) * a_new + d_pred).astype(F32)


newmark_2dof = newmark   # reference dyn.py:55-62
newmark_3dof = newmark


def get_markers_coords_2dof(x0, y0, d):
    """Reference: dyn.py:79-81."""
    d = f32(d)
    return (f32(x0) + d[0]).astype(F32), (f32(y0) + d[1]).astype(F32)


def get_markers_coords_3dof(x0, y0, xc, yc, d):
    """Reference: dyn.py:95-100."""
    d = f32(d)
    xr, yr = f32(x0) - F32(xc), f32(y0) - F32(yc)
    c, s = np.cos(d[2]), np.sin(d[2])
    return ((F32(xc) + d[0] + xr * c - yr * s).astype(F32),
            (F32(yc) + d[1] + xr * s + yr * c).astype(F32))


def get_markers_velocity_3dof(xm, ym, xc, yc, d, v):
    """Reference: dyn.py:115-120."""
    d, v = f32(d), f32(v)
    xr = f32(xm) - F32(xc) - d[0]
    yr = f32(ym) - F32(yc) - d[1]
    return np.stack([v[0] - v[2] * yr, v[1] + v[2] * xr], axis=-1).astype(F32)


def get_force_to_obj(h_markers):
    """Reference: dyn.py:136."""
    return f32(h_markers).sum(axis=0, dtype=F32)


def get_torque_to_obj(xm, ym, xc, yc, d, h_markers):
    """Reference: dyn.py:151-154."""
    d, h = f32(d), f32(h_markers)
    xr = f32(xm) - (F32(xc) + d[0])
    yr = f32(ym) - (F32(yc) + d[1])
    return np.sum(xr * h[:, 1] - yr * h[:, 0], dtype=F32)
