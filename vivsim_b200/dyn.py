"""Rigid-body dynamics on the HOST with the names of the reference's ``vivsim.dyn``
(north_star keeps the ODE on the host; a device-resident variant is vsb_body_newmark).
Scalars and small NumPy arrays, fp32."""

import numpy as np

_F = np.float32


def _a(x):
    return np.asarray(x, dtype=_F)


def newmark(a, v, d, h, m, k, c, dt=1, gamma=0.5, beta=0.25):
    """One Newmark-beta step; scalar or matrix (m, k, c)   (dyn.py:5-51)."""
    a, v, d, h = _a(a), _a(v), _a(d), _a(h)
    c1, c2 = gamma * dt, beta * dt ** 2
    v1 = v + _F(dt * (1 - gamma)) * a
    v2 = d + _F(dt) * v + _F(dt ** 2 * (0.5 - beta)) * a
    if np.ndim(m) > 0:
        m, k, c = _a(m), _a(k), _a(c)
        a_next = np.linalg.solve(m + _F(c1) * c + _F(c2) * k, h - c @ v1 - k @ v2).astype(_F)
    else:
        a_next = ((h - _F(c) * v1 - _F(k) * v2) / _F(m + c1 * c + c2 * k)).astype(_F)
    return a_next, (_F(c1) * a_next + v1).astype(_F), (_F(c2) * a_next + v2).astype(_F)


newmark_2dof = newmark   # dyn.py:55-62
newmark_3dof = newmark


def get_markers_coords_2dof(x_markers_init, y_markers_init, d):
    """dyn.py:68-81."""
    d = _a(d)
    return (_a(x_markers_init) + d[0]).astype(_F), (_a(y_markers_init) + d[1]).astype(_F)


def get_markers_coords_3dof(x_markers_init, y_markers_init, x_center_init, y_center_init, d):
    """dyn.py:84-100."""
    d = _a(d)
    xr, yr = _a(x_markers_init) - _F(x_center_init), _a(y_markers_init) - _F(y_center_init)
    cs, sn = np.cos(d[2]), np.sin(d[2])
    return ((_F(x_center_init) + d[0] + xr * cs - yr * sn).astype(_F),
            (_F(y_center_init) + d[1] + xr * sn + yr * cs).astype(_F))


def get_markers_velocity_3dof(x_markers, y_markers, x_center_init, y_center_init, d, v):
    """dyn.py:103-120."""
    d, v = _a(d), _a(v)
    xr = _a(x_markers) - _F(x_center_init) - d[0]
    yr = _a(y_markers) - _F(y_center_init) - d[1]
    return np.stack([v[0] - v[2] * yr, v[1] + v[2] * xr], axis=-1).astype(_F)


def get_force_to_obj(h_markers):
    """dyn.py:126-136."""
    return _a(h_markers).sum(axis=0, dtype=_F)


def get_torque_to_obj(x_markers, y_markers, x_center_init, y_center_init, d, h_markers):
    """dyn.py:139-154."""
    d, h = _a(d), _a(h_markers)
    xr = _a(x_markers) - (_F(x_center_init) + d[0])
    yr = _a(y_markers) - (_F(y_center_init) + d[1])
    return np.sum(xr * h[:, 1] - yr * h[:, 0], dtype=_F)
