"""NumPy fp32 restatement of the reference's ``vivsim/post.py`` (TEST INFRASTRUCTURE ONLY: imported by tests/,
never by the product).  Pinned against tests/golden/post.npz, which make_golden.py produces by executing the
unmodified reference source on oracle/jaxshim."""

import numpy as np

F32 = np.float32


def _grad(a, axis):
    """jnp.gradient(a, axis=axis), unit spacing, first-order edges (post.py:17-29 via jax.numpy.gradient)."""
    a = np.asarray(a, dtype=F32)
    n = a.shape[axis]
    if n < 2:
        raise ValueError("Shape of array too small to calculate a numerical gradient")
    out = np.empty_like(a)
    sl = [slice(None)] * a.ndim

    def at(s):
        t = list(sl)
        t[axis] = s
        return tuple(t)
    out[at(slice(1, -1))] = (a[at(slice(2, None))] - a[at(slice(None, -2))]) * F32(0.5)
    out[at(0)] = a[at(1)] - a[at(0)]
    out[at(-1)] = a[at(-1)] - a[at(-2)]
    return out


def velocity_magnitude(u):            # post.py:6-14
    u = np.asarray(u, dtype=F32)
    return np.sqrt(np.sum(u * u, axis=0, dtype=F32)).astype(F32)


def velocity_gradient(u):             # post.py:17-29
    u = np.asarray(u, dtype=F32)
    d = u.shape[0]
    return np.stack([np.stack([_grad(u[i], j) for j in range(d)]) for i in range(d)])


def vorticity(u):                     # post.py:32-55
    u = np.asarray(u, dtype=F32)
    if u.shape[0] == 2:
        return _grad(u[1], 0) - _grad(u[0], 1)
    return np.stack([_grad(u[2], 1) - _grad(u[1], 2), _grad(u[0], 2) - _grad(u[2], 0), _grad(u[1], 0) - _grad(u[0], 1)])


def vorticity_magnitude(u):           # post.py:58-66
    w = vorticity(u)
    return np.abs(w) if np.asarray(u).shape[0] == 2 else velocity_magnitude(w)


def divergence(u):                    # post.py:70-80
    u = np.asarray(u, dtype=F32)
    out = np.zeros(u.shape[1:], dtype=F32)
    for i in range(u.shape[0]):
        out = out + _grad(u[i], i)
    return out


def strain_rate(u):                   # post.py:83-92
    G = velocity_gradient(u)
    return (F32(0.5) * (G + np.swapaxes(G, 0, 1))).astype(F32)


def strain_rate_magnitude(u):         # post.py:95-103
    S = strain_rate(u)
    return np.sqrt(np.sum(S * S, axis=(0, 1), dtype=F32)).astype(F32)


def kinetic_energy(u):                # post.py:106-114
    u = np.asarray(u, dtype=F32)
    return (F32(0.5) * np.sum(u * u, axis=0, dtype=F32)).astype(F32)


def mean_kinetic_energy(u):           # post.py:117-126
    return F32(np.mean(kinetic_energy(u), dtype=np.float64))


def pressure(rho, cs2=1.0 / 3.0):     # post.py:129-139
    return (np.asarray(rho, dtype=F32) * F32(cs2)).astype(F32)


def enstrophy(u):                     # post.py:142-152
    w = vorticity(u)
    if np.asarray(u).shape[0] == 2:
        return (F32(0.5) * w * w).astype(F32)
    return (F32(0.5) * np.sum(w * w, axis=0, dtype=F32)).astype(F32)


def mean_enstrophy(u):                # post.py:155-160
    return F32(np.mean(enstrophy(u), dtype=np.float64))


def q_criterion(u):                   # post.py:163-177
    G = velocity_gradient(u)
    return (F32(-0.5) * np.einsum("ij...,ji...->...", G, G)).astype(F32)
