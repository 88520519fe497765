"""D3Q19 oracle with the reference's ``vivsim.lbm3d`` names (test infrastructure only)."""

import numpy as np

from . import core
from .core import F32, f32
from .lattice import D3Q19 as L
from .lbm import get_omega, get_velocity_correction, _characteristic  # noqa: F401 (same formulas, lbm3d/basic.py:149-167)


def _moment_basis():
    """Rows of the reference's non-orthogonal D3Q19 moment basis.
    Reference: lbm3d/collision/mrt.py:7-30 -- rebuilt here from polynomials in c
    and verified entry-for-entry against the literal matrix by the golden test."""
    c = L.c.astype(np.float64)
    x, y, z = c[:, 0], c[:, 1], c[:, 2]
    one = np.ones(19)
    rows = [one, x, y, z,
            x * x + y * y + z * z,
            2 * x * x - y * y - z * z,
            y * y - z * z,
            x * y, x * z, y * z,
            x * x * y, x * x * z, x * y * y, y * y * z, x * z * z, y * z * z,
            x * x * y * y, x * x * z * z, y * y * z * z]
    return np.array(rows)


M = _moment_basis()


def mrt_rates(omega):
    """diag(S).  Reference: lbm3d/collision/mrt.py:50-72."""
    return [0, 0, 0, 0, 1.1] + [omega] * 5 + [1.2] * 6 + [1.4] * 3


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
    """3-D KBC: shear part = full second-order Hermite projection of fneq.
    Reference: lbm3d/collision/kbc.py:29-42 (differs from the 2-D split)."""
    f = f32(f); feq = f32(feq)
    return core.kbc_from_split(f, feq, core.second_order_projection(L, f - feq), omega)


def boundary_equilibrium(f, loc, rho_wall=1, ux_wall=0, uy_wall=0, uz_wall=0):
    return core.boundary_equilibrium(L, f, loc, rho_wall, (ux_wall, uy_wall, uz_wall))


def boundary_nee(f, loc, rho_wall=1, ux_wall=0, uy_wall=0, uz_wall=0):
    return core.boundary_nee(L, f, loc, rho_wall, (ux_wall, uy_wall, uz_wall))


def boundary_nebb(f, loc, rho_wall=1, ux_wall=0, uy_wall=0, uz_wall=0):
    """f_in <- f_opp(in) + feq_in(rho_w,u_w) - feq_opp(in)(rho_w,u_w); no transverse
    correction.  Reference: lbm3d/boundary/nebb.py:16-32."""
    face = L.face(loc)
    f = f32(f).copy()
    rho, u = core.wall_state(L, f, face, rho_wall, (ux_wall, uy_wall, uz_wall))
    feq = core.equilibrium(L, rho, u)
    sel = core._idx(L, face.axis, face.wall)
    wall = f[sel].copy()
    new = {i: wall[L.opp[i]] + feq[i] - feq[L.opp[i]] for i in face.in_dirs}
    for i, v in new.items():
        wall[i] = v
    f[sel] = wall
    return f


def _velocity_from_pressure(f, face, rho_wall):
    """Reference: lbm3d/boundary/_helpers.py:66-78 (both tangential components
    copied from the neighbour layer's macroscopic velocity)."""
    f = f32(f)
    _, u_nb = core.macroscopic(L, f[core._idx(L, face.axis, face.neighbor)])
    rho = np.broadcast_to(f32(rho_wall), core.face_shape(L, f, face))
    vel = [u_nb[0], u_nb[1], u_nb[2]]
    vel[face.axis] = F32(face.sign) * (F32(1) - core.rho_wall_numerator(L, f, face) / rho)
    return vel


def _wrap(core_fn):
    def velocity(f, loc, ux_wall=0, uy_wall=0, uz_wall=0):
        rho = core.rho_from_velocity(L, f, L.face(loc), (ux_wall, uy_wall, uz_wall))
        return core_fn(f, loc, rho_wall=rho, ux_wall=ux_wall, uy_wall=uy_wall, uz_wall=uz_wall)

    def pressure(f, loc, rho_wall=1):
        ux, uy, uz = _velocity_from_pressure(f, L.face(loc), rho_wall)
        return core_fn(f, loc, rho_wall=rho_wall, ux_wall=ux, uy_wall=uy, uz_wall=uz)

    def force_corrected(f, loc, rho_wall=1, ux_wall=0, uy_wall=0, uz_wall=0,
                        gx_wall=0, gy_wall=0, gz_wall=0):
        ux, uy, uz = core.corrected_wall_velocity(
            (ux_wall, uy_wall, uz_wall), rho_wall, (gx_wall, gy_wall, gz_wall))
        return core_fn(f, loc, rho_wall=rho_wall, ux_wall=ux, uy_wall=uy, uz_wall=uz)

    return velocity, pressure, force_corrected


boundary_velocity_nee, boundary_pressure_nee, boundary_force_corrected_nee = _wrap(boundary_nee)
boundary_velocity_nebb, boundary_pressure_nebb, boundary_force_corrected_nebb = _wrap(boundary_nebb)
(boundary_velocity_equilibrium, boundary_pressure_equilibrium,
 boundary_force_corrected_equilibrium) = _wrap(boundary_equilibrium)


def boundary_bounce_back(f_before_stream, f, loc, ux_wall=0, uy_wall=0, uz_wall=0):
    """in_k <- pre-stream out_k + 2 w rho_w (c_in.u_w)/cs^2, rho_w = sum_q f_pre at the
    wall node; out_k is the MIRROR image of in_k (reference's out_dirs order).
    Reference: lbm3d/boundary/bb.py:9-34."""
    face = L.face(loc)
    f = f32(f).copy()
    sel = core._idx(L, face.axis, face.wall)
    pre = f32(f_before_stream)[sel]
    rho = pre.sum(axis=0, dtype=F32)
    _, u = core.wall_state(L, f, face, 1, (ux_wall, uy_wall, uz_wall))
    wall = f[sel].copy()
    for i, o in zip(face.in_dirs, face.out_dirs):
        cu = sum(F32(L.c[i, a]) * u[a] for a in range(3))
        wall[i] = pre[o] + F32(2) * L.w[i] * rho * cu / F32(1 / 3)
    f[sel] = wall
    return f


def boundary_specular_reflection(f_before_stream, f, loc, ux_wall=0, uy_wall=0, uz_wall=0):
    """Byte-for-byte the bounce-back body in the reference.  Reference: lbm3d/boundary/bb.py:37-53."""
    return boundary_bounce_back(f_before_stream, f, loc, ux_wall, uy_wall, uz_wall)


def boundary_characteristic(rho, u, loc="right"):
    """Reference: lbm3d/boundary/cbc.py:16-51."""
    return _characteristic(L, rho, u, loc)
