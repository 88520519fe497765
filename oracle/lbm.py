"""D2Q9 oracle with the reference's ``vivsim.lbm`` names (test infrastructure only)."""

import numpy as np

from . import core
from .core import F32, f32
from .lattice import D2Q9 as L

# reference lbm/collision/mrt.py:10-22
M = np.array([
    [1, 1, 1, 1, 1, 1, 1, 1, 1],
    [-4, -1, -1, -1, -1, 2, 2, 2, 2],
    [4, -2, -2, -2, -2, 1, 1, 1, 1],
    [0, 1, 0, -1, 0, 1, -1, -1, 1],
    [0, -2, 0, 2, 0, 1, -1, -1, 1],
    [0, 0, 1, 0, -1, 1, 1, -1, -1],
    [0, 0, -2, 0, 2, 1, 1, -1, -1],
    [0, 1, -1, 1, -1, 0, 0, 0, 0],
    [0, 0, 0, 0, 0, 1, -1, 1, -1]], dtype=np.float64)


def mrt_rates(omega):
    """diag(S).  Reference: lbm/collision/mrt.py:44."""
    return [0, 1.4, 1.4, 0, 1.2, 0, 1.2, omega, omega]


def get_omega(nu):
    """Reference: lbm/basic.py:172."""
    return 1 / (3 * nu + 0.5)


def get_velocity_correction(g, rho=1):
    """g / (2 rho).  Reference: lbm/basic.py:194."""
    return (f32(g) * F32(0.5) / f32(rho)).astype(F32)


def streaming(f): return core.streaming(L, f)
def get_macroscopic(f): return core.macroscopic(L, f)
def get_equilibrium(rho, u): return core.equilibrium(L, rho, u)
def collision_bgk(f, feq, omega): return core.collision_bgk(f, feq, omega)
def collision_mrt(f, feq, op): return core.collision_mrt(f, feq, op)
def collision_reg(f, feq, omega): return core.collision_reg(L, f, feq, omega)
def get_mrt_collision_operator(omega): return core.mrt_operator(M, mrt_rates(omega))
def get_mrt_forcing_operator(omega): return core.mrt_operator(M, mrt_rates(omega), forcing=True)
def get_guo_forcing_term(g, u): return core.guo_term(L, g, u)
def forcing_edm(f, g, u): return core.forcing_edm(L, f, g, u)
def forcing_guo_bgk(f, g, u, omega): return core.forcing_guo_bgk(L, f, g, u, omega)
def forcing_guo_mrt(f, g, u, fop): return core.forcing_guo_mrt(L, f, g, u, fop)
def obstacle_bounce_back(f, mask): return core.obstacle_bounce_back(L, f, mask)


def collision_kbc(f, feq, omega):
    """2-D KBC: the shear part keeps only the deviatoric moments
    N = Pxx - Pyy and Pxy of fneq.  Reference: lbm/collision/kbc.py:37-44."""
    f = f32(f); feq = f32(feq)
    fneq = f - feq
    n = fneq[1] - fneq[2] + fneq[3] - fneq[4]
    pxy = fneq[5] - fneq[6] + fneq[7] - fneq[8]
    shear = np.zeros_like(fneq)
    for k, s in enumerate((1, -1, 1, -1)):
        shear[1 + k] = F32(s) * n / F32(4)
        shear[5 + k] = F32(s) * pxy / F32(4)
    return core.kbc_from_split(f, feq, shear, omega)


# ------------------------------------------------------------------ boundaries
def boundary_equilibrium(f, loc, rho_wall=1, ux_wall=0, uy_wall=0):
    return core.boundary_equilibrium(L, f, loc, rho_wall, (ux_wall, uy_wall))


def boundary_nee(f, loc, rho_wall=1, ux_wall=0, uy_wall=0):
    return core.boundary_nee(L, f, loc, rho_wall, (ux_wall, uy_wall))


def boundary_nebb(f, loc, rho_wall=1, ux_wall=0, uy_wall=0):
    """Zou/He with transverse-momentum correction.  Reference: lbm/boundary/nebb.py:41-58."""
    face = L.face(loc)
    f = f32(f).copy()
    rho, u = core.wall_state(L, f, face, rho_wall, (ux_wall, uy_wall))
    un = F32(face.sign) * u[face.axis]
    ut = F32(face.sign) * u[1 - face.axis]
    sel = core._idx(L, face.axis, face.wall)
    wall = f[sel].copy()
    t0, t1 = face.tan_dirs
    shear = F32(0.5) * (wall[t0] - wall[t1]) * F32(face.sign)
    normal = F32(1 / 6) * un * rho
    tang = F32(0.5) * ut * rho
    i0, i1, i2 = face.in_dirs
    o0, o1, o2 = face.out_dirs
    wall[i0] = wall[o0] + F32(2 / 3) * un * rho
    wall[i1] = wall[o1] - shear + normal + tang
    wall[i2] = wall[o2] + shear + normal - tang
    f[sel] = wall
    return f


def _velocity_from_pressure(f, face, rho_wall):
    """Normal velocity from the prescribed density, tangential velocity copied
    from the adjacent fluid line.  Reference: lbm/boundary/_helpers.py:98-132."""
    f = f32(f)
    nb = f[core._idx(L, face.axis, face.neighbor)]
    rho_nb = nb.sum(axis=0, dtype=F32)
    un = F32(face.sign) * (F32(1) - core.rho_wall_numerator(L, f, face) / f32(rho_wall))
    ps, ns = face.pos_side_dirs, face.neg_side_dirs
    ut = np.zeros_like(rho_nb)
    for p, n in zip(_ref_order(face, ps), _ref_order(face, ns)):
        ut = ut + (nb[p] - nb[n])
    ut = ut / rho_nb
    vel = [None, None]
    vel[face.axis] = un
    vel[1 - face.axis] = ut
    return vel


def _ref_order(face, dirs):
    return dirs  # summation order only changes the last ulp


def _wrap(core_fn):
    def velocity(f, loc, ux_wall=0, uy_wall=0):
        face = L.face(loc)
        rho = core.rho_from_velocity(L, f, face, (ux_wall, uy_wall))
        return core_fn(f, loc, rho_wall=rho, ux_wall=ux_wall, uy_wall=uy_wall)

    def pressure(f, loc, rho_wall=1):
        ux, uy = _velocity_from_pressure(f, L.face(loc), rho_wall)
        return core_fn(f, loc, rho_wall=rho_wall, ux_wall=ux, uy_wall=uy)

    def force_corrected(f, loc, rho_wall=1, ux_wall=0, uy_wall=0, gx_wall=0, gy_wall=0):
        ux, uy = core.corrected_wall_velocity((ux_wall, uy_wall), rho_wall, (gx_wall, gy_wall))
        return core_fn(f, loc, rho_wall=rho_wall, ux_wall=ux, uy_wall=uy)

    return velocity, pressure, force_corrected


# reference lbm/boundary/{nee.py:63-65, nebb.py:60-62, eq.py:59-61}
boundary_velocity_nee, boundary_pressure_nee, boundary_force_corrected_nee = _wrap(boundary_nee)
boundary_velocity_nebb, boundary_pressure_nebb, boundary_force_corrected_nebb = _wrap(boundary_nebb)
(boundary_velocity_equilibrium, boundary_pressure_equilibrium,
 boundary_force_corrected_equilibrium) = _wrap(boundary_equilibrium)


def _reflect(f_before_stream, f, loc, ux_wall, uy_wall, swap):
    face = L.face(loc)
    f = f32(f).copy()
    pre = f32(f_before_stream)[core._idx(L, face.axis, face.wall)]
    u = (f32(ux_wall), f32(uy_wall))
    un = F32(face.sign) * u[face.axis]
    ut = F32(face.sign) * u[1 - face.axis]
    o0, o1, o2 = face.out_dirs
    vals = (pre[o0] + F32(2 / 3) * un, pre[o1] + F32(1 / 6) * (un + ut), pre[o2] + F32(1 / 6) * (un - ut))
    i0, i1, i2 = face.in_dirs
    targets = (i0, i2, i1) if swap else (i0, i1, i2)
    sel = core._idx(L, face.axis, face.wall)
    wall = f[sel].copy()
    for t, v in zip(targets, vals):
        wall[t] = v
    f[sel] = wall
    return f


def boundary_bounce_back(f_before_stream, f, loc, ux_wall=0, uy_wall=0):
    """in_k <- pre-stream out_k + momentum of a moving wall (rho = 1 assumed).
    Reference: lbm/boundary/bb.py:43-53."""
    return _reflect(f_before_stream, f, loc, ux_wall, uy_wall, swap=False)


def boundary_specular_reflection(f_before_stream, f, loc, ux_wall=0, uy_wall=0):
    """Same values, diagonal targets swapped.  Reference: lbm/boundary/bb.py:82-95."""
    return _reflect(f_before_stream, f, loc, ux_wall, uy_wall, swap=True)


def boundary_characteristic(rho, u, loc="right"):
    """Non-reflective characteristic update of (rho, u) on a face.
    Reference: lbm/boundary/cbc.py:14-53."""
    return _characteristic(L, rho, u, loc)


def _characteristic(lat, rho, u, loc):
    if loc not in lat.faces:
        raise ValueError("loc must name a face of the lattice")
    face = lat.face(loc)
    rho = f32(rho); u = f32(u)
    s = face.sign
    ks = (0, 1, 2) if s > 0 else (-1, -2, -3)
    take = lambda a, k: np.take(a, k, axis=face.axis)
    r1, r2, r3 = (take(rho, k) for k in ks)
    n1, n2, n3 = (take(u[face.axis], k) for k in ks)
    cs = F32(1 / np.sqrt(3))
    coef = F32(-0.5 * s)
    drho = coef * (F32(3) * r1 - F32(4) * r2 + r3)
    dun = coef * (F32(3) * n1 - F32(4) * n2 + n3)
    l_out = (n1 - F32(s) * cs) * (dun - F32(s) * cs / r1 * drho)
    vel = [None] * lat.d
    vel[face.axis] = n1 - F32(0.5) * l_out
    for a in face.tangential_axes:
        vel[a] = take(u[a], ks[0])
    return (r1 - F32(0.5) * r1 / cs * l_out).astype(F32), np.stack(vel).astype(F32)
