"""Known-answer cases shared by the CPU (oracle) and GPU (Stepper) analytic tests: the configurations the reference's
examples compare with an analytic solution.  Each case is (spec, f0, steps, check); check(f)
asserts on the final populations F_n (NumPy, reference convention)."""

import numpy as np

from oracle import lbm, lbm3d

F32 = np.float32


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def taylor_green(prepared):
    """examples/2d/taylor_green_vortex.py:28-75: u(t) = u(0) exp(-nu k^2 t), energy ~ exp(-2 nu k^2 t).
    prepared = False initialises like the example (rho = 1: the missing pressure field rings acoustically, ~1 % of |u|),
    True adds the exact solution's pressure field rho = 1 + 3 p."""
    n, u0, nu, steps = 64, 0.02, 0.02, 600
    k = 2 * np.pi / n
    x = np.arange(n, dtype=F32)[:, None]
    y = np.arange(n, dtype=F32)[None, :]

    def exact(t):
        d = np.exp(-nu * 2 * k * k * t)
        return np.stack([u0 * np.sin(k * x) * np.cos(k * y) * d, -u0 * np.cos(k * x) * np.sin(k * y) * d]).astype(F32)

    p0 = (-u0 * u0 / 4.0 * (np.cos(2 * k * x) + np.cos(2 * k * y))).astype(F32)
    rho0 = (1.0 - 3.0 * p0).astype(F32) if prepared else np.ones((n, n), F32)
    tol = 3e-3 if prepared else 1.5e-2
    spec = dict(dim=2, shape=(n, n), collision="bgk", omega=lbm.get_omega(nu), forcing=None, post=[])

    def check(f):
        rho, u = lbm.get_macroscopic(f)
        assert rel_l2(u, exact(steps)) < tol
        e0, e1 = np.mean(exact(0) ** 2), np.mean(u ** 2)
        assert abs(e1 / e0 / np.exp(-2 * nu * 2 * k * k * steps) - 1) < 6e-3
        assert abs(float(rho.mean()) - float(rho0.mean())) < 2e-5      # mass is conserved (to fp32 accumulation)

    return spec, lbm.get_equilibrium(rho0, exact(0)), steps, check


def couette():
    """examples/2d/couette_flow.py:51-91: BGK + NEE walls, linear profile U0 y / (NY - 1)."""
    n, u0, nu, steps = 16, 0.08, 0.1, 6000
    spec = dict(dim=2, shape=(n, n), collision="bgk", omega=lbm.get_omega(nu), forcing=None,
                post=[("nee", "bottom", {}), ("nee", "top", {"ux_wall": u0})])

    def check(f):
        _, u = lbm.get_macroscopic(f)
        exact = np.broadcast_to(u0 * np.arange(n, dtype=F32) / (n - 1), (n, n))
        assert rel_l2(u[0], exact) < 1e-3
        assert np.abs(u[1]).max() < 1e-5

    return spec, lbm.get_equilibrium(np.ones((n, n), F32), np.zeros((2, n, n), F32)), steps, check


POISEUILLE_KINDS = ["bgk_edm", "bgk_guo", "mrt_guo", "kbc_edm"]


def poiseuille(kind):
    """examples/2d/poiseuille_channel.py:80-148,176-181: the four collision / forcing recipes against
    g / (2 nu) y (H - y), H = NY - 1, with the half-force velocity correction applied to the output."""
    nx = ny = 10
    nu, gx, steps = 0.2, 1e-3, 3000                      # 3000 steps = 7 viscous times H^2 / nu
    coll, forcing = kind.split("_")
    spec = dict(dim=2, shape=(nx, ny), collision=coll, omega=lbm.get_omega(nu), forcing=forcing, g=(gx, 0.0),
                post=[("force_corrected_nebb", "top", {"gx_wall": gx}), ("force_corrected_nebb", "bottom", {"gx_wall": gx})])
    u_init = np.zeros((2, nx, ny), F32)
    u_init[0] = -gx / 2

    def check(f):
        rho, u = lbm.get_macroscopic(f)
        ux = u[0] + gx / (2 * rho)
        yy = np.arange(ny, dtype=F32)
        exact = gx / (2 * nu) * yy * (ny - 1 - yy)
        assert np.abs(ux[nx // 2] - exact).max() < 2e-3 * exact.max()
        assert np.abs(ux - ux[0]).max() < 1e-5           # uniform along the periodic direction

    return spec, lbm.get_equilibrium(np.ones((nx, ny), F32), u_init), steps, check


def abc_flow():
    """examples/3d/abc_flow.py:59-68,128-130: regularised collision, velocity decays as exp(-nu k^2 t)."""
    n, u0, nu, steps = 32, 0.02, 0.005, 200
    k = 2 * np.pi / n
    x, y, z = np.meshgrid(*(np.arange(n, dtype=F32),) * 3, indexing="ij")
    u_init = np.stack([u0 * np.sin(k * z) + u0 * np.cos(k * y), u0 * np.sin(k * x) + u0 * np.cos(k * z),
                       u0 * np.sin(k * y) + u0 * np.cos(k * x)]).astype(F32)
    spec = dict(dim=3, shape=(n, n, n), collision="reg", omega=lbm3d.get_omega(nu), forcing=None, post=[])

    def check(f):
        _, u = lbm3d.get_macroscopic(f)
        assert rel_l2(u, u_init * np.exp(-nu * k * k * steps)) < 2e-2

    return spec, lbm3d.get_equilibrium(np.ones((n, n, n), F32), u_init), steps, check
