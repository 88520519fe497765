"""Field diagnostics with the names and signatures of the reference's ``vivsim.post``
(vivsim/post.py:1-211), one fused sm_100a kernel per call behind the C ABI (``vsb_post_field`` /
``vsb_post_mean``): the gradient stencil and the tensor contraction happen in registers, so a diagnostic costs
one read of ``u`` and one write of the result instead of the reference's chain of ``jnp.gradient`` / ``stack`` /
``einsum`` temporaries.

Arrays are fp32 torch CUDA tensors: u (2, NX, NY) or (3, NX, NY, NZ), rho (NX, NY[, NZ])."""

import ctypes as C

import torch

from . import _lib as L


def _u(u):
    u = L.dev(u, name="u")
    if u.ndim not in (3, 4) or u.shape[0] != u.ndim - 1:
        raise ValueError(f"u must have shape (2, NX, NY) or (3, NX, NY, NZ), got {tuple(u.shape)}")
    return u


def _field(kind, u, lead=()):
    u = _u(u)
    spatial = tuple(u.shape[1:])
    if kind not in ("velocity_magnitude", "kinetic_energy") and min(spatial) < 2:
        # numpy / jax.numpy.gradient: "Shape of array too small to calculate a numerical gradient"
        raise ValueError("Shape of array too small to calculate a numerical gradient, "
                         "at least 2 elements are required along every axis.")
    out = torch.empty(tuple(lead) + spatial, dtype=torch.float32, device=u.device)
    grid = L.grid_of(spatial)
    L.check(L.lib().vsb_post_field(C.byref(grid), L.DIAG[kind], L.ptr(u), C.c_float(0.0), L.ptr(out), L.stream()))
    return out


def _mean(kind, u):
    u = _u(u)
    spatial = tuple(u.shape[1:])
    if kind != "kinetic_energy" and min(spatial) < 2:
        raise ValueError("Shape of array too small to calculate a numerical gradient, "
                         "at least 2 elements are required along every axis.")
    ws = torch.empty(1, dtype=torch.float64, device=u.device)
    out = torch.empty((), dtype=torch.float32, device=u.device)
    grid = L.grid_of(spatial)
    L.check(L.lib().vsb_post_mean(C.byref(grid), L.DIAG[kind], L.ptr(u), C.c_float(0.0), L.ptr(ws), L.ptr(out),
                                  L.stream()))
    return out


def velocity_magnitude(u):
    """|u|   (post.py:6-14)."""
    return _field("velocity_magnitude", u)


def velocity_gradient(u):
    """G[i, j] = du_i / dx_j, shape (dim, dim, *spatial)   (post.py:17-29)."""
    d = _u(u).shape[0]
    return _field("velocity_gradient", u, (d, d))


def vorticity(u):
    """Scalar dv/dx - du/dy in 2-D, curl vector (3, *spatial) in 3-D   (post.py:32-55)."""
    return _field("vorticity", u, () if _u(u).shape[0] == 2 else (3,))


def vorticity_magnitude(u):
    """post.py:58-66."""
    return _field("vorticity_magnitude", u)


def divergence(u):
    """post.py:70-80."""
    return _field("divergence", u)


def strain_rate(u):
    """0.5 (G + G^T), shape (dim, dim, *spatial)   (post.py:83-92)."""
    d = _u(u).shape[0]
    return _field("strain_rate", u, (d, d))


def strain_rate_magnitude(u):
    """Frobenius norm of the strain-rate tensor   (post.py:95-103)."""
    return _field("strain_rate_magnitude", u)


def kinetic_energy(u):
    """0.5 |u|^2   (post.py:106-114)."""
    return _field("kinetic_energy", u)


def mean_kinetic_energy(u):
    """Domain mean of the kinetic energy, 0-d tensor   (post.py:117-126).  The field is not materialised."""
    return _mean("kinetic_energy", u)


def pressure(rho, cs2=1.0 / 3.0):
    """rho * cs2   (post.py:129-139)."""
    rho = L.dev(rho, name="rho")
    if rho.ndim not in (2, 3):
        raise ValueError(f"rho must have shape (NX, NY) or (NX, NY, NZ), got {tuple(rho.shape)}")
    out = torch.empty_like(rho)
    grid = L.grid_of(rho.shape)
    L.check(L.lib().vsb_post_field(C.byref(grid), L.DIAG["pressure"], L.ptr(rho), C.c_float(cs2), L.ptr(out),
                                   L.stream()))
    return out


def enstrophy(u):
    """0.5 |omega|^2   (post.py:142-152)."""
    return _field("enstrophy", u)


def mean_enstrophy(u):
    """Domain mean of the enstrophy, 0-d tensor   (post.py:155-160)."""
    return _mean("enstrophy", u)


def q_criterion(u):
    """-0.5 G_ij G_ji   (post.py:163-177)."""
    return _field("q_criterion", u)


# Deprecated aliases of the reference (post.py:180-211)
def calculate_curl(u):
    return vorticity(u)


def calculate_vorticity(u):
    return vorticity(u)


def calculate_vorticity_dimensionless(u, l, u0):
    return vorticity(u) * l / u0


def calculate_velocity_magnitude(u):
    return velocity_magnitude(u)
