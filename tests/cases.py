"""Step descriptions (``spec`` dicts) for the composed recipes in tests/golden/recipes.npz.

The same spec is fed to ``oracle.recipes`` (CPU) and to ``vivsim_b200.Stepper`` (GPU)."""

import numpy as np

from oracle import lbm, lbm3d
from oracle.core import F32


def _eq2(nx, ny, ux=0.0, uy=0.0):
    u = np.zeros((2, nx, ny), dtype=F32); u[0] = ux; u[1] = uy
    return lbm.get_equilibrium(np.ones((nx, ny), dtype=F32), u)


def cavity(g):
    nx, ny, u0, nu = g["cavity_params"]; nx, ny = int(nx), int(ny)
    spec = dict(dim=2, shape=(nx, ny), collision="bgk", omega=lbm.get_omega(nu), forcing=None,
                post=[("nee", "left", {}), ("nee", "right", {}), ("nee", "bottom", {}),
                      ("nee", "top", {"ux_wall": float(u0)})])
    return spec, _eq2(nx, ny), 30, "cavity_f30"


def cavity_kbc_topfirst(g):
    nx, ny, u0, nu = g["cavity_params"]; nx, ny = int(nx), int(ny)
    spec = dict(dim=2, shape=(nx, ny), collision="kbc", omega=lbm.get_omega(nu), forcing=None,
                post=[("nee", "top", {"ux_wall": float(u0)}), ("nee", "left", {}), ("nee", "right", {}),
                      ("nee", "bottom", {})])
    return spec, _eq2(nx, ny), 20, "cavity_kbc_topfirst_f20"


def poiseuille(g, kind):
    nx, ny, gx, nu = g["pois_params"]; nx, ny = int(nx), int(ny)
    coll, forcing = kind.split("_")
    spec = dict(dim=2, shape=(nx, ny), collision=coll, omega=lbm.get_omega(nu), forcing=forcing,
                g=(float(gx), 0.0),
                post=[("force_corrected_nebb", "top", {"gx_wall": float(gx)}),
                      ("force_corrected_nebb", "bottom", {"gx_wall": float(gx)})])
    f0 = _eq2(nx, ny, ux=-float(F32(gx) * F32(0.5)))
    return spec, f0, 40, f"pois_{kind}_f40"


def cylinder(g, recipe):
    nx, ny, u0, nu, rad, x0, y0, size, ds = g["cyl_params"]
    nx, ny, x0, y0, size = int(nx), int(ny), int(x0), int(y0), int(size)
    markers = np.stack([g["cyl_mx"], g["cyl_my"]], axis=1)
    post = [("force_corrected_nebb", "left", {"ux_wall": float(u0)}), ("equilibrium", "right", {"ux_wall": float(u0)})]
    if recipe == "kbc_edm":
        spec = dict(dim=2, shape=(nx, ny), collision="kbc", omega=lbm.get_omega(nu), forcing="edm",
                    ib=dict(markers=markers, ds=float(ds), kernel="peskin4", n_iter=3, window=((x0, y0), (size, size))),
                    post=post)
        return spec, _eq2(nx, ny, ux=float(u0)), 25, "cyl_f25"
    from oracle import ib
    spec = dict(dim=2, shape=(nx, ny), collision="bgk", omega=lbm.get_omega(nu), forcing="guo",
                ib=dict(markers=markers, ds=ib.get_ds(markers), kernel="peskin4", n_iter=5, window=((x0, y0), (size, size))),
                post=post)
    return spec, _eq2(nx, ny, ux=float(u0)), 10, "c2_f10"


def viv(g):
    (nx, ny, D, u0, nu, Mm, K, C, area, X0, Y0, size, mds, vy0) = g["viv_params"]
    nx, ny, X0, Y0, size = int(nx), int(ny), int(X0), int(Y0), int(size)
    markers = np.stack([g["viv_MX"], g["viv_MY"]], axis=1)
    spec = dict(dim=2, shape=(nx, ny), collision="reg", omega=lbm.get_omega(nu), forcing="edm",
                ib=dict(markers=markers, ds=float(mds), kernel="peskin4", n_iter=1, window=((X0, Y0), (size, size))),
                post=[("force_corrected_nebb", "left", {"ux_wall": float(u0)}),
                      ("equilibrium", "right", {"ux_wall": float(u0)})])
    body = dict(m=float(Mm), k=float(K), c=float(C), added_mass=float(area))
    state = (np.zeros(2, F32), np.array([0, vy0], F32), np.zeros(2, F32))
    return spec, body, _eq2(nx, ny, ux=float(u0)), state, 20


def viv_rotation(g):
    """Elastically mounted ellipse with three degrees of freedom (tests/golden/make_golden.py rotation_fixture)."""
    (nx, ny, u0, nu, xc, yc, X0, Y0, size, th0, vy0, om0) = g["rot_params"]
    nx, ny, X0, Y0, size = int(nx), int(ny), int(X0), int(Y0), int(size)
    markers = np.stack([g["rot_MX"], g["rot_MY"]], axis=1)
    spec = dict(dim=2, shape=(nx, ny), collision="bgk", omega=lbm.get_omega(nu), forcing="edm",
                ib=dict(markers=markers, ds=g["rot_ds"], kernel="peskin4", n_iter=3, window=((X0, Y0), (size, size))),
                post=[("force_corrected_nebb", "left", {"ux_wall": float(u0)}),
                      ("equilibrium", "right", {"ux_wall": float(u0)})])
    body = dict(m=g["rot_M"], k=g["rot_K"], c=g["rot_C"], added_mass=g["rot_added"], center=(float(xc), float(yc)),
                rotation=True, n_dof=3)
    state = (np.array([0, 0, th0], F32), np.array([0, vy0, om0], F32), np.zeros(3, F32))
    return spec, body, _eq2(nx, ny, ux=float(u0)), state, 30


def text_mask(g):
    nx, ny, u0, nu = g["text_params"]; nx, ny = int(nx), int(ny)
    spec = dict(dim=2, shape=(nx, ny), collision="kbc", omega=lbm.get_omega(nu), forcing=None,
                post=[("nee", "bottom", {"uy_wall": float(u0)}), ("equilibrium", "top", {"uy_wall": float(u0)}),
                      ("mask", g["text_mask"])])
    return spec, _eq2(nx, ny, uy=float(u0)), 15, "text_f15"


def sphere(g):
    p = g["sphere_params"]
    shape = tuple(int(x) for x in p[:3]); u0, nu = float(p[3]), float(p[4])
    o = tuple(int(x) for x in p[6:9]); size = tuple(int(x) for x in p[9:12])
    from oracle import ib3d
    verts, faces = g["sphere_verts"], g["sphere_faces"]
    spec = dict(dim=3, shape=shape, collision="kbc", omega=lbm3d.get_omega(nu), forcing="edm",
                ib=dict(markers=verts, ds=ib3d.get_ds(verts, faces), kernel="peskin4", n_iter=3, window=(o, size)),
                post=[("nebb", "left", {"ux_wall": u0}), ("equilibrium", "right", {"ux_wall": u0})])
    u = np.zeros((3,) + shape, dtype=F32); u[0] = u0
    return spec, lbm3d.get_equilibrium(np.ones(shape, dtype=F32), u), 10, "sphere_f10"


def mrt3(g):
    p = g["mrt3_params"]
    shape = tuple(int(x) for x in p[:3]); omega = float(p[3])
    spec = dict(dim=3, shape=shape, collision="mrt", omega=omega, forcing="guo",
                g=tuple(float(x) for x in p[4:7]), post=[])
    return spec, lbm3d.get_equilibrium(np.ones(shape, dtype=F32), g["mrt3_u_init"]), 12, "mrt3_f12"


def all_fluid_cases(g):
    yield "cavity", cavity(g)
    yield "cavity_kbc_topfirst", cavity_kbc_topfirst(g)
    for kind in ("bgk_edm", "bgk_guo", "mrt_guo", "kbc_edm", "reg_edm"):
        yield f"poiseuille_{kind}", poiseuille(g, kind)
    yield "cylinder_kbc_edm", cylinder(g, "kbc_edm")
    yield "cylinder_c2", cylinder(g, "c2")
    yield "text_mask", text_mask(g)
    yield "sphere", sphere(g)
    yield "mrt3", mrt3(g)
