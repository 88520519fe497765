"""Generate tests/golden/*.npz by executing the UNMODIFIED reference source.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

jax/jaxlib are absent from this image, so the reference (pure Python over
jax.numpy) is executed on the NumPy-backed stand-in in ``oracle/jaxshim``
(float64->float32 demotion after every primitive, functional ``.at`` updates).
The fixtures therefore pin the reference's algorithm, indexing, operation order
and constants -- not XLA's fp32 rounding.  The committed .npz files travel to the
GPU box; /root/reference does not.
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("VIVSIM_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle", "jaxshim"))
sys.path.insert(0, REF)

import jax  # noqa: E402  (the shim)
import jax.numpy as jnp  # noqa: E402
from vivsim import dyn, ib, ib3d, lbm, lbm3d  # noqa: E402
from vivsim.lbm.lattice import D2Q9  # noqa: E402
from vivsim.lbm3d.lattice import D3Q19  # noqa: E402
from vivsim.lbm.collision import mrt as mrt2  # noqa: E402
from vivsim.lbm3d.collision import mrt as mrt3, reg as reg3  # noqa: E402

assert jax.__version__.endswith("numpy-shim")
F32 = np.float32


def A(x):
    return np.asarray(x)


def J(x):
    return jnp.asarray(np.array(x, copy=True))


def save(name, d):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: A(v) for k, v in d.items()})
    print(f"{name}.npz: {len(d)} arrays, {os.path.getsize(path) / 1024:.1f} KiB")


def perturbed_state(mod, shape, dim, rng, amp=0.05):
    rho = (1 + amp * rng.standard_normal(shape)).astype(F32)
    u = (amp * rng.standard_normal((dim,) + shape)).astype(F32)
    feq = A(mod.get_equilibrium(J(rho), J(u)))
    f = (feq * (1 + 0.02 * rng.standard_normal(feq.shape))).astype(F32)
    return rho, u, feq, f


# ---------------------------------------------------------------- lattice tables
def lattice_fixture():
    out = {}
    for tag, lat in (("d2q9", D2Q9), ("d3q19", D3Q19)):
        out[f"{tag}_c"], out[f"{tag}_w"], out[f"{tag}_opp"] = lat.c, lat.w, lat.opp_dirs
        for loc, spec in lat.boundary_spec.items():
            out[f"{tag}_{loc}_in"] = spec.in_dirs
            out[f"{tag}_{loc}_out"] = spec.out_dirs
            out[f"{tag}_{loc}_sign_axis"] = np.array([spec.normal_sign, spec.normal_axis])
            if tag == "d2q9":
                out[f"{tag}_{loc}_tan"] = spec.tan_dirs
                out[f"{tag}_{loc}_pos"] = spec.pos_side_dirs
                out[f"{tag}_{loc}_neg"] = spec.neg_side_dirs
            else:
                out[f"{tag}_{loc}_zero"] = spec.zero_dirs
    out["d2q9_M"], out["d3q19_M"] = mrt2.M, mrt3.M
    out["d3q19_P"] = A(reg3._SECOND_ORDER_PROJECTION)
    for om in (0.8, 1.7):
        out[f"d2q9_mrt_op_{om}"] = A(lbm.get_mrt_collision_operator(om))
        out[f"d2q9_mrt_fop_{om}"] = A(lbm.get_mrt_forcing_operator(om))
        out[f"d3q19_mrt_op_{om}"] = A(lbm3d.get_mrt_collision_operator(om))
        out[f"d3q19_mrt_fop_{om}"] = A(lbm3d.get_mrt_forcing_operator(om))
    save("lattice", out)


# ---------------------------------------------------------------- per-op fixtures
def ops_fixture(tag, mod, shape, locs, seed):
    dim = len(shape)
    rng = np.random.default_rng(seed)
    rho, u, feq, f = perturbed_state(mod, shape, dim, rng)
    g = (1e-3 * rng.standard_normal((dim,) + shape)).astype(F32)
    omega = 1.7
    mask = rng.random(shape) < 0.15
    f_pre = (f * (1 + 0.01 * rng.standard_normal(f.shape))).astype(F32)
    out = dict(rho=rho, u=u, feq=feq, f=f, g=g, omega=F32(omega), mask=mask, f_pre=f_pre)

    out["streaming"] = mod.streaming(J(f))
    r, uu = mod.get_macroscopic(J(f))
    out["macro_rho"], out["macro_u"] = r, uu
    # edge-slice form accepted by the reference (lbm/basic.py:100-101)
    r, uu = mod.get_macroscopic(J(f[:, 1]))
    out["macro_edge_rho"], out["macro_edge_u"] = r, uu
    out["equilibrium"] = mod.get_equilibrium(J(rho), J(u))
    out["bgk"] = mod.collision_bgk(J(f), J(feq), omega)
    out["kbc"] = mod.collision_kbc(J(f), J(feq), omega)
    out["reg"] = mod.collision_reg(J(f), J(feq), omega)
    op, fop = mod.get_mrt_collision_operator(omega), mod.get_mrt_forcing_operator(omega)
    out["mrt_op"], out["mrt_fop"] = op, fop
    out["mrt"] = mod.collision_mrt(J(f), J(feq), op)
    out["guo_term"] = mod.get_guo_forcing_term(J(g), J(u))
    out["edm"] = mod.forcing_edm(J(f), J(g), J(u))
    out["guo_bgk"] = mod.forcing_guo_bgk(J(f), J(g), J(u), omega)
    out["guo_mrt"] = mod.forcing_guo_mrt(J(f), J(g), J(u), fop)
    out["vel_corr"] = mod.get_velocity_correction(J(g), J(rho))
    out["obstacle_bb"] = mod.obstacle_bounce_back(J(f), J(mask))

    comps = ["ux_wall", "uy_wall", "uz_wall"][:dim]
    gcomps = ["gx_wall", "gy_wall", "gz_wall"][:dim]
    for loc in locs:
        spec = (D2Q9 if dim == 2 else D3Q19).boundary_spec[loc]
        fshape = tuple(n for a, n in enumerate(shape) if a != spec.normal_axis)
        uw_s = {c: 0.03 * (k + 1) * (-1) ** k for k, c in enumerate(comps)}
        gw_s = {c: 2e-3 * (k + 1) for k, c in enumerate(gcomps)}
        uw_a = {c: (0.04 * rng.standard_normal(fshape)).astype(F32) for c in comps}
        rw_a = (1 + 0.03 * rng.standard_normal(fshape)).astype(F32)
        for c in comps:
            out[f"{loc}_arr_{c}"] = uw_a[c]
        out[f"{loc}_arr_rho"] = rw_a
        out[f"{loc}_scalar_u"] = np.array([uw_s[c] for c in comps], dtype=F32)
        out[f"{loc}_scalar_g"] = np.array([gw_s[c] for c in gcomps], dtype=F32)
        uw_aj = {c: J(v) for c, v in uw_a.items()}
        for kind in ("nee", "nebb", "equilibrium"):
            core = getattr(mod, f"boundary_{kind}")
            out[f"{kind}_{loc}_default"] = core(J(f), loc)
            out[f"{kind}_{loc}_scalar"] = core(J(f), loc, rho_wall=1.02, **uw_s)
            out[f"{kind}_{loc}_array"] = core(J(f), loc, rho_wall=J(rw_a), **uw_aj)
            out[f"velocity_{kind}_{loc}_scalar"] = getattr(mod, f"boundary_velocity_{kind}")(J(f), loc, **uw_s)
            out[f"velocity_{kind}_{loc}_array"] = getattr(mod, f"boundary_velocity_{kind}")(J(f), loc, **uw_aj)
            out[f"pressure_{kind}_{loc}_scalar"] = getattr(mod, f"boundary_pressure_{kind}")(J(f), loc, rho_wall=0.98)
            out[f"force_corrected_{kind}_{loc}_scalar"] = getattr(mod, f"boundary_force_corrected_{kind}")(
                J(f), loc, rho_wall=1.01, **uw_s, **gw_s)
        out[f"bounce_back_{loc}_default"] = mod.boundary_bounce_back(J(f_pre), J(f), loc)
        out[f"bounce_back_{loc}_scalar"] = mod.boundary_bounce_back(J(f_pre), J(f), loc, **uw_s)
        out[f"specular_{loc}_scalar"] = mod.boundary_specular_reflection(J(f_pre), J(f), loc, **uw_s)
        r, uu = mod.boundary_characteristic(J(rho), J(u), loc)
        out[f"cbc_{loc}_rho"], out[f"cbc_{loc}_u"] = r, uu
    save(tag, out)


# ---------------------------------------------------------------- immersed boundary
def ib_fixture():
    rng = np.random.default_rng(7)
    out = {}
    r = np.linspace(-2.6, 2.6, 105).astype(F32)
    out["r"] = r
    out["peskin3"], out["peskin4"], out["cosine4"] = (
        ib.kernel_peskin_3pt(J(r)), ib.kernel_peskin_4pt(J(r)), ib.kernel_cosine_4pt(J(r)))

    # 2-D: closed curve of markers well inside a 24 x 20 grid
    nx, ny, m = 24, 20, 40
    th = np.linspace(0, 2 * np.pi, m, endpoint=False)
    mx = (11.3 + 5.2 * np.cos(th)).astype(F32)
    my = (9.6 + 4.1 * np.sin(th)).astype(F32)
    coords = np.stack([mx, my], axis=1)
    u = (0.05 * rng.standard_normal((2, nx, ny))).astype(F32)
    tgt = (0.02 * rng.standard_normal((m, 2))).astype(F32)
    vals = (0.1 * rng.standard_normal((m, 2))).astype(F32)
    out.update(mx=mx, my=my, u2=u, tgt2=tgt, vals2=vals, shape2=np.array([nx, ny]))
    out["ds2_closed"], out["ds2_open"], out["area2"] = ib.get_ds(J(coords)), ib.get_ds(J(coords), closed=False), ib.get_area(J(coords))
    for kname, kern in (("peskin4", ib.kernel_peskin_4pt), ("peskin3", ib.kernel_peskin_3pt),
                        ("cosine4", ib.kernel_cosine_4pt)):
        w, idx = ib.get_ib_stencil(J(mx), J(my), ny, kernel=kern)
        out[f"w2_{kname}"], out[f"idx2_{kname}"] = w, idx
    w, idx = ib.get_ib_stencil(J(mx), J(my), ny)
    out["interp2"] = ib.interpolate(J(u), w, idx)
    out["spread2"] = ib.spread(J(vals), J(u), w, idx)
    ds = out["ds2_closed"]
    for n_iter in (1, 5):
        gg, hh = ib.multi_direct_forcing(J(u), w, idx, J(tgt), ds, n_iter=n_iter)
        out[f"mdf2_g_{n_iter}"], out[f"mdf2_h_{n_iter}"] = gg, hh
    gg, hh = ib.multi_direct_forcing(J(u), w, idx, J(tgt), 0.7, n_iter=3)
    out["mdf2_g_scalar_ds"], out["mdf2_h_scalar_ds"] = gg, hh

    # 3-D: icosphere (162 vertices) inside 16 x 14 x 12, built by this script (not the reference)
    from oracle.ib3d import icosphere
    shape = (16, 14, 12)
    verts, faces = icosphere(3.7, (7.4, 6.8, 5.9), 2)
    u3 = (0.05 * rng.standard_normal((3,) + shape)).astype(F32)
    tgt3 = (0.02 * rng.standard_normal((verts.shape[0], 3))).astype(F32)
    vals3 = (0.1 * rng.standard_normal((verts.shape[0], 3))).astype(F32)
    out.update(verts=verts, faces=faces, u3=u3, tgt3=tgt3, vals3=vals3, shape3=np.array(shape))
    out["tri_areas"], out["surf_area"], out["volume"], out["ds3"] = (
        ib3d.get_triangle_areas(J(verts), J(faces)), ib3d.get_surface_area(J(verts), J(faces)),
        ib3d.get_volume(J(verts), J(faces)), ib3d.get_ds(J(verts), J(faces)))
    w3, idx3 = ib3d.get_ib_stencil(J(verts), shape)
    out["w3"], out["idx3"] = w3, idx3
    out["interp3"] = ib3d.interpolate(J(u3), w3, idx3)
    out["spread3"] = ib3d.spread(J(vals3), J(u3), w3, idx3)
    gg, hh = ib3d.multi_direct_forcing(J(u3), w3, idx3, J(tgt3), out["ds3"], n_iter=3)
    out["mdf3_g"], out["mdf3_h"] = gg, hh
    save("ib", out)


# ---------------------------------------------------------------- dynamics
def dyn_fixture():
    out = {}
    a, v, d, h = (np.array(x, dtype=F32) for x in ([0.01, -0.02], [0.1, 0.05], [0.3, -0.1], [0.7, 0.2]))
    out.update(a=a, v=v, d=d, h=h)
    res = dyn.newmark_2dof(J(a), J(v), J(d), J(h), 31.4, 0.8, 0.05)
    out["nm_scalar_a"], out["nm_scalar_v"], out["nm_scalar_d"] = res
    m = np.diag([10.0, 12.0]).astype(F32); k = np.array([[2.0, 0.3], [0.3, 1.5]], dtype=F32); c = (0.1 * np.eye(2)).astype(F32)
    out.update(m=m, k=k, c=c)
    res = dyn.newmark(J(a), J(v), J(d), J(h), J(m), J(k), J(c))
    out["nm_matrix_a"], out["nm_matrix_v"], out["nm_matrix_d"] = res
    x0 = np.linspace(3, 9, 11).astype(F32); y0 = np.linspace(-2, 4, 11).astype(F32)
    d3 = np.array([0.4, -0.2, 0.3], dtype=F32); v3 = np.array([0.05, 0.02, -0.01], dtype=F32)
    hm = np.stack([np.sin(x0), np.cos(y0)], axis=1).astype(F32)
    out.update(x0=x0, y0=y0, d3=d3, v3=v3, hm=hm)
    out["c2x"], out["c2y"] = dyn.get_markers_coords_2dof(J(x0), J(y0), J(d))
    xm, ym = dyn.get_markers_coords_3dof(J(x0), J(y0), 6.0, 1.0, J(d3))
    out["c3x"], out["c3y"] = xm, ym
    out["v3m"] = dyn.get_markers_velocity_3dof(xm, ym, 6.0, 1.0, J(d3), J(v3))
    out["force"] = dyn.get_force_to_obj(J(hm))
    out["torque"] = dyn.get_torque_to_obj(xm, ym, 6.0, 1.0, J(d3), J(hm))
    save("dyn", out)


# ---------------------------------------------------------------- composed steps
def recipes_fixture():
    out = {}

    # (1) README cavity (README.md:104-122) at 24 x 20, 30 steps
    nx, ny, u0 = 24, 20, 0.3
    omega = lbm.get_omega(0.1)
    f = lbm.get_equilibrium(jnp.ones((nx, ny)), jnp.zeros((2, nx, ny)))
    for _ in range(30):
        rho, u = lbm.get_macroscopic(f)
        f = lbm.collision_bgk(f, lbm.get_equilibrium(rho, u), omega)
        f = lbm.streaming(f)
        f = lbm.boundary_nee(f, loc="left")
        f = lbm.boundary_nee(f, loc="right")
        f = lbm.boundary_nee(f, loc="bottom")
        f = lbm.boundary_nee(f, loc="top", ux_wall=u0)
    out["cavity_f30"] = f
    out["cavity_params"] = np.array([nx, ny, u0, 0.1])

    # (1b) cavity order of lid_driven_cavity.py:55-64 (top first) with KBC
    f = lbm.get_equilibrium(jnp.ones((nx, ny)), jnp.zeros((2, nx, ny)))
    for _ in range(20):
        rho, u = lbm.get_macroscopic(f)
        f = lbm.collision_kbc(f, lbm.get_equilibrium(rho, u), omega)
        f = lbm.streaming(f)
        f = lbm.boundary_nee(f, loc="top", ux_wall=u0)
        f = lbm.boundary_nee(f, loc="left")
        f = lbm.boundary_nee(f, loc="right")
        f = lbm.boundary_nee(f, loc="bottom")
    out["cavity_kbc_topfirst_f20"] = f

    # (2) Poiseuille recipes (poiseuille_channel.py:80-148) at 12 x 10, 40 steps
    nx, ny, gx, nu = 12, 10, 1e-3, 0.2
    omega = lbm.get_omega(nu)
    g = jnp.zeros((2, nx, ny)).at[0].set(gx)
    mop, mfop = lbm.get_mrt_collision_operator(omega), lbm.get_mrt_forcing_operator(omega)
    out["pois_params"] = np.array([nx, ny, gx, nu])

    def pois(kind):
        f = lbm.get_equilibrium(jnp.ones((nx, ny)), jnp.zeros((2, nx, ny)) - lbm.get_velocity_correction(g))
        for _ in range(40):
            rho, u = lbm.get_macroscopic(f)
            if kind in ("bgk_guo", "mrt_guo"):
                u = u + lbm.get_velocity_correction(g, rho)
            feq = lbm.get_equilibrium(rho, u)
            if kind == "bgk_edm":
                f = lbm.forcing_edm(lbm.collision_bgk(f, feq, omega), g, u)
            elif kind == "bgk_guo":
                f = lbm.forcing_guo_bgk(lbm.collision_bgk(f, feq, omega), g, u, omega)
            elif kind == "mrt_guo":
                f = lbm.forcing_guo_mrt(lbm.collision_mrt(f, feq, mop), g, u, mfop)
            elif kind == "kbc_edm":
                f = lbm.forcing_edm(lbm.collision_kbc(f, feq, omega), g, u)
            elif kind == "reg_edm":
                f = lbm.forcing_edm(lbm.collision_reg(f, feq, omega), g, u)
            f = lbm.streaming(f)
            f = lbm.boundary_force_corrected_nebb(f, loc="top", gx_wall=gx)
            f = lbm.boundary_force_corrected_nebb(f, loc="bottom", gx_wall=gx)
        return f

    for kind in ("bgk_edm", "bgk_guo", "mrt_guo", "kbc_edm", "reg_edm"):
        out[f"pois_{kind}_f40"] = pois(kind)

    # (3) fixed cylinder, IB window + MDF + EDM (flow_pass_cylinder.py:97-125) at 48 x 32, 25 steps
    nx, ny, u0, nu, m, rad = 48, 32, 0.08, 0.02, 48, 5.0
    omega = lbm.get_omega(nu)
    th = np.linspace(0, 2 * np.pi, m, endpoint=False)
    mx = (16.3 + rad * np.cos(th)).astype(F32); my = (15.7 + rad * np.sin(th)).astype(F32)
    x0, y0, size = 8, 8, 17
    ds = 2 * np.pi * rad / m
    w, idx = ib.get_ib_stencil(J(mx - x0), J(my - y0), size)
    f = lbm.get_equilibrium(jnp.ones((nx, ny)), jnp.zeros((2, nx, ny)).at[0].set(u0))
    hs = []
    for _ in range(25):
        rho, u = lbm.get_macroscopic(f)
        f = lbm.collision_kbc(f, lbm.get_equilibrium(rho, u), omega)
        ib_u = jax.lax.dynamic_slice(u, (0, x0, y0), (2, size, size))
        ib_f = jax.lax.dynamic_slice(f, (0, x0, y0), (9, size, size))
        ib_g, hm = ib.multi_direct_forcing(ib_u, w, idx, jnp.zeros((m, 2)), ds, n_iter=3)
        hs.append(A(dyn.get_force_to_obj(hm)))
        f = jax.lax.dynamic_update_slice(f, lbm.forcing_edm(ib_f, ib_g, ib_u), (0, x0, y0))
        f = lbm.streaming(f)
        f = lbm.boundary_force_corrected_nebb(f, loc="left", ux_wall=u0)
        f = lbm.boundary_equilibrium(f, loc="right", ux_wall=u0)
    out["cyl_f25"], out["cyl_h"] = f, np.array(hs)
    out["cyl_mx"], out["cyl_my"] = mx, my
    out["cyl_params"] = np.array([nx, ny, u0, nu, rad, x0, y0, size, ds])

    # (3b) SURVEY C2 recipe: BGK + MDF(5) on u + Guo shift + guo_bgk (benchmark.py fixtures), 48 x 32, 10 steps
    f = lbm.get_equilibrium(jnp.ones((nx, ny)), jnp.zeros((2, nx, ny)).at[0].set(u0))
    mds = ib.get_ds(J(np.stack([mx, my], axis=1)))
    for _ in range(10):
        rho, u = lbm.get_macroscopic(f)
        ib_u = jax.lax.dynamic_slice(u, (0, x0, y0), (2, size, size))
        ib_g, hm = ib.multi_direct_forcing(ib_u, w, idx, jnp.zeros((m, 2)), mds, n_iter=5)
        g = jax.lax.dynamic_update_slice(jnp.zeros((2, nx, ny)), ib_g, (0, x0, y0))
        u = u + lbm.get_velocity_correction(g, rho)
        f = lbm.collision_bgk(f, lbm.get_equilibrium(rho, u), omega)
        f = lbm.forcing_guo_bgk(f, g, u, omega)
        f = lbm.streaming(f)
        f = lbm.boundary_force_corrected_nebb(f, loc="left", ux_wall=u0)
        f = lbm.boundary_equilibrium(f, loc="right", ux_wall=u0)
    out["c2_f10"], out["c2_h_last"] = f, hm

    # (4) VIV moving cylinder (vortex_induced_vibration.py:96-148) at 64 x 40, 20 steps
    nx, ny, D, u0 = 64, 40, 8, 0.06
    nu = u0 * D / 100
    omega = lbm.get_omega(nu)
    m = 4 * D
    th = np.linspace(0, 2 * np.pi, m, endpoint=False)
    MX = (20 + 0.5 * D * np.cos(th)).astype(F32); MY = (20 + 0.5 * D * np.sin(th)).astype(F32)
    area = np.pi * (D / 2) ** 2
    MR, UR = 10, 5
    fn = u0 / (UR * D); Mm = area * MR; K = (2 * np.pi * fn) ** 2 * Mm * (1 + 1 / MR); C = 0.0
    pad = 4
    X0 = int(20 - 0.5 * D - pad); Y0 = int(20 - 0.5 * D - pad); size = D + 2 * pad
    mds = 2 * np.pi * (D / 2) / m
    f = lbm.get_equilibrium(jnp.ones((nx, ny)), jnp.zeros((2, nx, ny)).at[0].set(u0))
    d = jnp.zeros(2); v = jnp.zeros(2).at[1].set(0.3 * u0); a = jnp.zeros(2)
    dh = []
    for _ in range(20):
        rho, u = lbm.get_macroscopic(f)
        f = lbm.collision_reg(f, lbm.get_equilibrium(rho, u), omega)
        ibx = (X0 + d[0]).astype(jnp.int32); iby = (Y0 + d[1]).astype(jnp.int32)
        ib_u = jax.lax.dynamic_slice(u, (0, ibx, iby), (2, size, size))
        ib_f = jax.lax.dynamic_slice(f, (0, ibx, iby), (9, size, size))
        a_old, v_old, d_old = a, v, d
        mxx, myy = dyn.get_markers_coords_2dof(J(MX), J(MY), d)
        w, idx = ib.get_ib_stencil(mxx - ibx, myy - iby, size, kernel=ib.kernel_peskin_4pt, stencil_radius=2)
        mv = jnp.repeat(v[None, :], m, axis=0)
        ib_g, hm = ib.multi_direct_forcing(ib_u, w, idx, mv, mds, n_iter=1)
        h = dyn.get_force_to_obj(hm)
        h += a * area
        a, v, d = dyn.newmark_2dof(a_old, v_old, d_old, h, Mm, K, C)
        f = jax.lax.dynamic_update_slice(f, lbm.forcing_edm(ib_f, ib_g, ib_u), (0, ibx, iby))
        f = lbm.streaming(f)
        f = lbm.boundary_force_corrected_nebb(f, loc="left", ux_wall=u0)
        f = lbm.boundary_equilibrium(f, loc="right", ux_wall=u0)
        dh.append(np.concatenate([A(d), A(v), A(a), A(h)]))
    out["viv_f20"], out["viv_dvah"] = f, np.array(dh)
    out["viv_MX"], out["viv_MY"] = MX, MY
    out["viv_params"] = np.array([nx, ny, D, u0, nu, Mm, K, C, area, X0, Y0, size, mds, 0.3 * u0])

    # (5) obstacle mask after BCs (flow_through_text.py:70-79) at 20 x 24, 15 steps
    nx, ny, u0 = 20, 24, 0.05
    omega = lbm.get_omega(0.02)
    rng = np.random.default_rng(3)
    mask = np.zeros((nx, ny), dtype=bool); mask[6:9, 8:14] = True; mask[13, 5:9] = True; mask[0, :3] = True
    f = lbm.get_equilibrium(jnp.ones((nx, ny)), jnp.zeros((2, nx, ny)).at[1].set(u0))
    for _ in range(15):
        rho, u = lbm.get_macroscopic(f)
        f = lbm.collision_kbc(f, lbm.get_equilibrium(rho, u), omega)
        f = lbm.streaming(f)
        f = lbm.boundary_nee(f, loc="bottom", uy_wall=u0)
        f = lbm.boundary_equilibrium(f, loc="top", uy_wall=u0)
        f = lbm.obstacle_bounce_back(f, J(mask))
    out["text_f15"], out["text_mask"] = f, mask
    out["text_params"] = np.array([nx, ny, u0, 0.02])

    # (6) 3-D fixed sphere (flow_past_sphere.py:139-171) at 28 x 16 x 16, 10 steps
    from oracle.ib3d import icosphere
    shape = (28, 16, 16); u0 = 0.05; Dd = 6.0
    nu = u0 * Dd / 200
    omega = lbm3d.get_omega(nu)
    verts, faces = icosphere(Dd / 2, (9.3, 7.8, 8.2), 2)
    o = (3, 2, 2); size = (13, 12, 12)
    w3, idx3 = ib3d.get_ib_stencil(J(verts - np.array(o, dtype=F32)), size)
    ds3 = ib3d.get_ds(J(verts), J(faces))
    f = lbm3d.get_equilibrium(jnp.ones(shape), jnp.zeros((3,) + shape).at[0].set(u0))
    for _ in range(10):
        rho, u = lbm3d.get_macroscopic(f)
        f = lbm3d.collision_kbc(f, lbm3d.get_equilibrium(rho, u), omega)
        ib_u = jax.lax.dynamic_slice(u, (0,) + o, (3,) + size)
        ib_f = jax.lax.dynamic_slice(f, (0,) + o, (19,) + size)
        ib_g, hm = ib3d.multi_direct_forcing(ib_u, w3, idx3, jnp.zeros((verts.shape[0], 3)), ds3, n_iter=3)
        f = jax.lax.dynamic_update_slice(f, lbm3d.forcing_edm(ib_f, ib_g, ib_u), (0,) + o)
        f = lbm3d.streaming(f)
        f = lbm3d.boundary_nebb(f, loc="left", ux_wall=u0)
        f = lbm3d.boundary_equilibrium(f, loc="right", ux_wall=u0)
    out["sphere_f10"], out["sphere_h_last"] = f, hm
    out["sphere_verts"], out["sphere_faces"] = verts, faces
    out["sphere_params"] = np.array(list(shape) + [u0, nu, Dd] + list(o) + list(size), dtype=np.float64)

    # (7) 3-D MRT + Guo body force, periodic (C5 collision/forcing pair), 10 x 8 x 6, 12 steps
    shape = (10, 8, 6)
    omega = 1.3
    rng = np.random.default_rng(11)
    g = jnp.zeros((3,) + shape).at[0].set(2e-4).at[2].set(-1e-4)
    u_init = (0.03 * rng.standard_normal((3,) + shape)).astype(F32)
    f = lbm3d.get_equilibrium(jnp.ones(shape), J(u_init))
    mop, mfop = lbm3d.get_mrt_collision_operator(omega), lbm3d.get_mrt_forcing_operator(omega)
    for _ in range(12):
        rho, u = lbm3d.get_macroscopic(f)
        u = u + lbm3d.get_velocity_correction(g, rho)
        f = lbm3d.collision_mrt(f, lbm3d.get_equilibrium(rho, u), mop)
        f = lbm3d.forcing_guo_mrt(f, g, u, mfop)
        f = lbm3d.streaming(f)
    out["mrt3_f12"], out["mrt3_u_init"] = f, u_init
    out["mrt3_params"] = np.array(list(shape) + [omega, 2e-4, 0.0, -1e-4])
    save("recipes", out)


# ---------------------------------------------------------------- 3-DOF rigid body (translation + rotation)
def rotation_fixture():
    """An elastically mounted ELLIPSE with three degrees of freedom (x, y, rotation), composed from the reference's
    own functions the way examples/2d/vortex_induced_vibration.py:96-148 composes the 2-DOF case: marker coordinates
    and velocities from dyn.get_markers_coords_3dof / get_markers_velocity_3dof (dyn.py:84-120), torque from
    dyn.get_torque_to_obj (dyn.py:139-154), matrix-form dyn.newmark_3dof (dyn.py:36-42).  72 x 48, 30 steps."""
    out = {}
    nx, ny, u0 = 72, 48, 0.06
    A_, B_ = 6.0, 3.5
    nu = u0 * 2 * A_ / 100
    omega = lbm.get_omega(nu)
    m = 40
    th = np.linspace(0, 2 * np.pi, m, endpoint=False)
    XC, YC = 24.3, 23.6
    MX = (XC + A_ * np.cos(th)).astype(F32); MY = (YC + B_ * np.sin(th)).astype(F32)
    area = np.pi * A_ * B_
    inertia = 0.25 * area * 10 * (A_ ** 2 + B_ ** 2)
    M = np.diag([10 * area, 10 * area, inertia]).astype(F32)
    K = np.array([[0.02, 0.0, 0.0], [0.0, 0.05, 0.01], [0.0, 0.01, 0.9]], dtype=F32)
    C = np.array([[0.01, 0.0, 0.0], [0.0, 0.02, 0.0], [0.0, 0.0, 0.3]], dtype=F32)
    added = np.array([area, area, 0.0], dtype=F32)
    pad = 5
    X0 = int(XC - A_ - pad); Y0 = int(YC - A_ - pad); size = int(2 * A_ + 2 * pad)
    mds = A(ib.get_ds(J(np.stack([MX, MY], axis=1))))
    f = lbm.get_equilibrium(jnp.ones((nx, ny)), jnp.zeros((2, nx, ny)).at[0].set(u0))
    d = jnp.zeros(3).at[2].set(0.35); v = jnp.zeros(3).at[1].set(0.3 * u0).at[2].set(0.004); a = jnp.zeros(3)
    rec = []
    for _ in range(30):
        rho, u = lbm.get_macroscopic(f)
        f = lbm.collision_bgk(f, lbm.get_equilibrium(rho, u), omega)
        ibx = (X0 + d[0]).astype(jnp.int32); iby = (Y0 + d[1]).astype(jnp.int32)
        ib_u = jax.lax.dynamic_slice(u, (0, ibx, iby), (2, size, size))
        ib_f = jax.lax.dynamic_slice(f, (0, ibx, iby), (9, size, size))
        mxx, myy = dyn.get_markers_coords_3dof(J(MX), J(MY), XC, YC, d)
        mv = dyn.get_markers_velocity_3dof(mxx, myy, XC, YC, d, v)
        w, idx = ib.get_ib_stencil(mxx - ibx, myy - iby, size, kernel=ib.kernel_peskin_4pt, stencil_radius=2)
        ib_g, hm = ib.multi_direct_forcing(ib_u, w, idx, mv, J(mds), n_iter=3)
        h = jnp.concatenate([dyn.get_force_to_obj(hm), dyn.get_torque_to_obj(mxx, myy, XC, YC, d, hm)[None]])
        h = h + a * J(added)
        a, v, d = dyn.newmark_3dof(a, v, d, h, J(M), J(K), J(C))
        f = jax.lax.dynamic_update_slice(f, lbm.forcing_edm(ib_f, ib_g, ib_u), (0, ibx, iby))
        f = lbm.streaming(f)
        f = lbm.boundary_force_corrected_nebb(f, loc="left", ux_wall=u0)
        f = lbm.boundary_equilibrium(f, loc="right", ux_wall=u0)
        rec.append(np.concatenate([A(d), A(v), A(a), A(h)]))
    out["rot_f30"], out["rot_dvah"] = f, np.array(rec)
    out["rot_MX"], out["rot_MY"], out["rot_ds"] = MX, MY, mds
    out["rot_M"], out["rot_K"], out["rot_C"], out["rot_added"] = M, K, C, added
    out["rot_params"] = np.array([nx, ny, u0, nu, XC, YC, X0, Y0, size, 0.35, 0.3 * u0, 0.004])
    save("rotation", out)


# ---------------------------------------------------------------- post.py diagnostics
def post_fixture():
    from vivsim import post
    out = {}
    rng = np.random.default_rng(21)
    for tag, shape in (("2d", (13, 10)), ("2d_thin", (2, 9)), ("3d", (7, 6, 5)), ("3d_thin", (2, 2, 3))):
        dim = len(shape)
        # smooth part + noise so that gradients are neither trivial nor pure noise
        grids = np.meshgrid(*[np.arange(n) for n in shape], indexing="ij")
        u = np.stack([0.05 * np.sin(0.7 * grids[d] + 0.3 * grids[(d + 1) % dim] + d) for d in range(dim)])
        u = (u + 0.01 * rng.standard_normal(u.shape)).astype(F32)
        rho = (1 + 0.02 * rng.standard_normal(shape)).astype(F32)
        out[f"{tag}_u"], out[f"{tag}_rho"] = u, rho
        for name in ("velocity_magnitude", "velocity_gradient", "vorticity", "vorticity_magnitude", "divergence",
                     "strain_rate", "strain_rate_magnitude", "kinetic_energy", "mean_kinetic_energy", "enstrophy",
                     "mean_enstrophy", "q_criterion", "calculate_curl", "calculate_vorticity",
                     "calculate_velocity_magnitude"):
            out[f"{tag}_{name}"] = getattr(post, name)(J(u))
        out[f"{tag}_pressure"] = post.pressure(J(rho))
        out[f"{tag}_pressure_cs2"] = post.pressure(J(rho), 0.25)
        out[f"{tag}_vorticity_dimensionless"] = post.calculate_vorticity_dimensionless(J(u), 20.0, 0.05)
    save("post", out)


# ---------------------------------------------------------------- multigrid.py transfers
def multigrid_fixture():
    from vivsim import multigrid
    out = {}
    rng = np.random.default_rng(22)
    # left/right: the edge line runs along y (fine ny = 2 coarse ny); up/down: along x
    for tag, fine_shape, coarse_shape in (("lr", (5, 12), (4, 6)), ("ud", (14, 6), (7, 3)), ("lr_min", (2, 2), (1, 1)),
                                          ("ud_min", (2, 2), (1, 1))):
        ff = rng.standard_normal((9,) + fine_shape).astype(F32)
        fc = rng.standard_normal((9,) + coarse_shape).astype(F32)
        out[f"{tag}_fine"], out[f"{tag}_coarse"] = ff, fc
        for d in (("left", "right") if tag.startswith("lr") else ("up", "down")):
            out[f"{tag}_f2c_{d}"] = multigrid.fine_to_coarse(J(ff), J(fc), d)
            out[f"{tag}_c2f_{d}"] = multigrid.coarse_to_fine(J(fc), J(ff), d)
        out[f"{tag}_f2c_other"] = multigrid.fine_to_coarse(J(ff), J(fc), "top")     # unknown dir: unchanged
    out["omega"] = np.array([[nu, lv, multigrid.get_omega(nu, lv)] for nu in (0.01, 0.1) for lv in (-1, 0, 1, 2)])
    out["coord"] = np.array([multigrid.coord_to_indices(13.5, 7.25, 4, 2, lv) for lv in (-1, 0, 1, 2)])
    for i, (w, h, lv, bx, by) in enumerate(((8, 6, 0, 0, 0), (8, 6, 1, 2, 0), (8, 6, -1, 0, 1))):
        f, rho, u = multigrid.init_grid(w, h, lv, bx, by)
        out[f"init{i}_args"] = np.array([w, h, lv, bx, by])
        out[f"init{i}_shapes"] = np.array(A(f).shape + A(rho).shape + A(u).shape)
        out[f"init{i}_sums"] = np.array([A(f).sum(), A(rho).mean(), A(u).sum()])
    save("multigrid", out)


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    np.seterr(invalid="ignore")
    which = sys.argv[1:]   # e.g. `make_golden.py post multigrid` regenerates only those files

    def want(name):
        return not which or name in which
    if want("lattice"):
        lattice_fixture()
    if want("ops2d"):
        ops_fixture("ops2d", lbm, (12, 8), ("left", "right", "top", "bottom"), seed=1)
    if want("ops3d"):
        ops_fixture("ops3d", lbm3d, (7, 6, 5), ("left", "right", "bottom", "top", "back", "front"), seed=2)
    if want("ib"):
        ib_fixture()
    if want("dyn"):
        dyn_fixture()
    if want("recipes"):
        recipes_fixture()
    if want("rotation"):
        rotation_fixture()
    if want("post"):
        post_fixture()
    if want("multigrid"):
        multigrid_fixture()
