"""Lattice-generic part of the oracle (NumPy fp32; test infrastructure only).

Each function takes a ``Lattice`` from ``oracle.lattice`` and follows the cited
reference lines.  Where the reference's 2-D and 3-D modules use different
formulas (KBC split, NEBB, bounce-back, pressure wrapper) the two variants are
kept separate in ``oracle/lbm.py`` and ``oracle/lbm3d.py``.
"""

import numpy as np

F32 = np.float32


def f32(x):
    return np.asarray(x, dtype=F32)


def _bcast(vec, ndim):
    """(Q,) table -> (Q,1,..,1) for broadcasting over ``ndim`` spatial axes."""
    return vec.reshape((-1,) + (1,) * ndim)


# ----------------------------------------------------------------- a1 streaming
def streaming(lat, f):
    """Periodic push ``f_q(x + c_q) <- f_q(x)``.

    Reference: lbm/basic.py:44-85, lbm3d/basic.py:28-87 (``shift_*_pos`` gives
    ``new[x] = old[x-1]`` = ``np.roll(+1)``; diagonals compose two shifts).
    """
    f = f32(f)
    out = np.empty_like(f)
    axes = tuple(range(lat.d))
    for q in range(lat.q):
        out[q] = np.roll(f[q], shift=tuple(int(s) for s in lat.c[q]), axis=axes)
    return out


# --------------------------------------------------------------- a2 macroscopic
def macroscopic(lat, f):
    """rho = sum_q f, u = sum_q c_q f / rho.

    Reference: lbm/basic.py:107-110 (fixed index sets), lbm3d/basic.py:102-105
    (einsum with the int32 velocity table).  Works on full fields and on edge
    slices ``(Q, N)``.
    """
    f = f32(f)
    rho = f.sum(axis=0, dtype=F32)
    u = np.zeros((lat.d,) + rho.shape, dtype=F32)
    for a in range(lat.d):
        pos = [q for q in range(lat.q) if lat.c[q, a] > 0]
        neg = [q for q in range(lat.q) if lat.c[q, a] < 0]
        mom = f[pos].sum(axis=0, dtype=F32) - f[neg].sum(axis=0, dtype=F32)
        u[a] = mom / rho
    return rho, u


# --------------------------------------------------------------- a3 equilibrium
def equilibrium(lat, rho, u):
    """feq_q = rho w_q (1 + 3 c.u + 4.5 (c.u)^2 - 1.5 u.u).

    Reference: lbm/basic.py:132-135, lbm3d/basic.py:121-130.
    """
    rho = f32(rho)
    u = f32(u)
    nd = rho.ndim
    cu = np.zeros((lat.q,) + rho.shape, dtype=F32)
    for a in range(lat.d):
        cu += _bcast(lat.c[:, a].astype(F32), nd) * u[a]
    usq = (u * u).sum(axis=0, dtype=F32)
    return (rho * _bcast(lat.w, nd) * (F32(1) + F32(3) * cu + F32(4.5) * cu * cu - F32(1.5) * usq)).astype(F32)


# ------------------------------------------------------------------ a4 BGK
def collision_bgk(f, feq, omega):
    """(1-omega) f + omega feq.  Reference: lbm/basic.py:156, lbm3d/basic.py:146."""
    return (F32(1 - omega) * f32(f) + F32(omega) * f32(feq)).astype(F32)


# ------------------------------------------------------------------ a5 MRT
def collision_mrt(f, feq, op):
    """f + A (feq - f) with A = M^-1 S M passed in.

    Reference: lbm/collision/mrt.py:88, lbm3d/collision/mrt.py:96-98.
    """
    f = f32(f)
    return (f + np.tensordot(f32(op), f32(feq) - f, axes=([1], [0]))).astype(F32)


def mrt_operator(M, s_diag, forcing=False):
    """M^-1 S M (collision) or M^-1 (I - S/2) M (forcing), as fp32.

    Reference: lbm/collision/mrt.py:47-63 + lbm/forcing/guo.py:62-79 and the
    lbm3d twins (mrt.py:76-80, guo.py:58-67).  Evaluated in float64 and rounded
    once; JAX evaluates the same products in fp32 (difference ~1e-7 relative).
    """
    M = np.asarray(M, dtype=np.float64)
    s = np.asarray(s_diag, dtype=np.float64)
    core = np.diag(1.0 - 0.5 * s) if forcing else np.diag(s)
    return (np.linalg.inv(M) @ core @ M).astype(F32)


# ---------------------------------------------------------- a7 second-order part
def second_order_projection(lat, fneq):
    """P fneq with P_qr = w_q/(2 cs^4) (c_q c_q - cs^2 I):(c_r c_r).

    Reference: closed form lbm/collision/reg.py:23-47; literal matrix
    lbm3d/collision/reg.py:10-41 (equal to this closed form).
    """
    fneq = f32(fneq)
    nd = fneq.ndim - 1
    out = np.zeros_like(fneq)
    for a in range(lat.d):
        for b in range(lat.d):
            cab = (lat.c[:, a] * lat.c[:, b]).astype(F32)
            pi_ab = (_bcast(cab, nd) * fneq).sum(axis=0, dtype=F32)
            q_ab = cab - (F32(1 / 3) if a == b else F32(0))
            out += _bcast(lat.w * F32(4.5) * q_ab, nd) * pi_ab
    return out


def collision_reg(lat, f, feq, omega):
    """feq + (1-omega) P (f-feq).  Reference: lbm/collision/reg.py:49, lbm3d/collision/reg.py:60-62."""
    f = f32(f); feq = f32(feq)
    return (feq + F32(1 - omega) * second_order_projection(lat, f - feq)).astype(F32)


def kbc_from_split(f, feq, shear, omega):
    """Entropic mixing common to both KBC variants.

    Reference: lbm/collision/kbc.py:47-59, lbm3d/collision/kbc.py:32-42.
    """
    f = f32(f); feq = f32(feq)
    high = (f - feq) - shear
    inv = F32(1) / (feq + F32(1e-20))
    sh = (high * shear * inv).sum(axis=0, dtype=F32)
    hh = (high * high * inv).sum(axis=0, dtype=F32)
    inv_w = F32(1.0 / omega)
    half_gamma = inv_w - (F32(1) - inv_w) * sh / (hh + F32(1e-20))
    return (f - F32(omega) * (shear + half_gamma * high)).astype(F32)


# ------------------------------------------------------------------ a8 / a9 forcing
def guo_term(lat, g, u):
    """G_q = w_q [3 (c_q - u).g + 9 (c_q.u)(c_q.g)].

    Reference: lbm/forcing/guo.py:21-33, lbm3d/forcing/guo.py:13-38.
    """
    g = f32(g); u = f32(u)
    nd = u.ndim - 1
    cu = np.zeros((lat.q,) + u.shape[1:], dtype=F32)
    cg = np.zeros_like(cu)
    for a in range(lat.d):
        ca = _bcast(lat.c[:, a].astype(F32), nd)
        cu += ca * u[a]
        cg += ca * g[a]
    ug = (u * g).sum(axis=0, dtype=F32)
    return (_bcast(lat.w, nd) * (F32(3) * (cg - ug) + F32(9) * cu * cg)).astype(F32)


def forcing_edm(lat, f, g, u):
    """f + G.  Reference: lbm/forcing/edm.py:31, lbm3d/forcing/edm.py:26."""
    return (f32(f) + guo_term(lat, g, u)).astype(F32)


def forcing_guo_bgk(lat, f, g, u, omega):
    """f + (1 - omega/2) G.  Reference: lbm/forcing/guo.py:56-57, lbm3d/forcing/guo.py:54-55."""
    return (f32(f) + guo_term(lat, g, u) * F32(1 - 0.5 * omega)).astype(F32)


def forcing_guo_mrt(lat, f, g, u, fop):
    """f + B G with B = M^-1 (I - S/2) M.  Reference: lbm/forcing/guo.py:106-107, lbm3d/forcing/guo.py:84-87."""
    G = guo_term(lat, g, u)
    return (f32(f) + np.tensordot(f32(fop), G, axes=([1], [0]))).astype(F32)


# ------------------------------------------------------------------ boundaries
def _idx(lat, axis, k):
    """Index tuple selecting layer ``k`` along spatial ``axis`` (after the Q axis)."""
    return (slice(None),) + tuple(k if a == axis else slice(None) for a in range(lat.d))


def face_shape(lat, f, face):
    return tuple(f.shape[1 + a] for a in range(lat.d) if a != face.axis)


def wall_state(lat, f, face, rho_wall, u_wall):
    """Broadcast scalars / arrays to the face.  Reference: lbm/boundary/_helpers.py:53-77,
    lbm3d/boundary/_helpers.py:19-33."""
    shape = face_shape(lat, f, face)
    rho = np.broadcast_to(f32(rho_wall), shape).astype(F32)
    u = np.stack([np.broadcast_to(f32(c), shape) for c in u_wall]).astype(F32)
    return rho, u


def boundary_equilibrium(lat, f, loc, rho_wall, u_wall):
    """wall <- feq(rho_w, u_w).  Reference: lbm/boundary/eq.py:45-56, lbm3d/boundary/eq.py:16-26."""
    face = lat.face(loc)
    f = f32(f).copy()
    rho, u = wall_state(lat, f, face, rho_wall, u_wall)
    f[_idx(lat, face.axis, face.wall)] = equilibrium(lat, rho, u)
    return f


def boundary_nee(lat, f, loc, rho_wall, u_wall):
    """wall <- feq(rho_w,u_w) + (f - feq(rho,u)) at the adjacent fluid layer.

    Reference: lbm/boundary/nee.py:43-60, lbm3d/boundary/nee.py:16-32.
    """
    face = lat.face(loc)
    f = f32(f).copy()
    rho, u = wall_state(lat, f, face, rho_wall, u_wall)
    f_nb = f[_idx(lat, face.axis, face.neighbor)]
    rho_nb, u_nb = macroscopic(lat, f_nb)
    f[_idx(lat, face.axis, face.wall)] = equilibrium(lat, rho, u) + (f_nb - equilibrium(lat, rho_nb, u_nb))
    return f


def rho_wall_numerator(lat, f, face):
    """sum_{zero} f + 2 sum_{out} f on the wall.  Reference: lbm/boundary/_helpers.py:135-145,
    lbm3d/boundary/_helpers.py:35-43."""
    wall = f32(f)[_idx(lat, face.axis, face.wall)]
    return wall[list(face.zero_dirs)].sum(axis=0, dtype=F32) + F32(2) * wall[list(face.out_dirs)].sum(axis=0, dtype=F32)


def rho_from_velocity(lat, f, face, u_wall):
    """rho_w = numerator / (1 - u_n), u_n = sign * u[axis].  Reference: lbm/boundary/_helpers.py:80-95,
    lbm3d/boundary/_helpers.py:58-63."""
    un = F32(face.sign) * f32(u_wall[face.axis])
    return (rho_wall_numerator(lat, f, face) / (F32(1) - un)).astype(F32)


def corrected_wall_velocity(u_wall, rho_wall, g_wall):
    """u_w - g_w / (2 rho_w).  Reference: lbm/boundary/_helpers.py:156-177, lbm3d/boundary/_helpers.py:81-88."""
    return tuple(f32(u) - f32(g) * F32(0.5) / f32(rho_wall) for u, g in zip(u_wall, g_wall))


def obstacle_bounce_back(lat, f, mask):
    """masked cells: f_q <- f_opp(q).  Reference: lbm/boundary/bb.py:110, lbm3d/boundary/bb.py:59."""
    f = f32(f).copy()
    m = np.asarray(mask, dtype=bool)
    f[:, m] = f[:, m][lat.opp]
    return f
