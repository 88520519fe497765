"""GPU: every per-function CUDA kernel (through the C ABI) against the committed golden fixtures
(reference source) and against the NumPy oracle on fresh seeded inputs."""

import numpy as np
import pytest
import torch

from conftest import assert_bitexact, assert_close
import oracle.lbm, oracle.lbm3d, oracle.ib, oracle.ib3d

pytestmark = pytest.mark.gpu

LOCS = {"ops2d": ("left", "right", "top", "bottom"), "ops3d": ("left", "right", "bottom", "top", "back", "front")}
COMPS = {"ops2d": ("ux_wall", "uy_wall"), "ops3d": ("ux_wall", "uy_wall", "uz_wall")}


def T(x):
    return torch.as_tensor(np.ascontiguousarray(x), device="cuda")


def N(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def mods(tag):
    import vivsim_b200 as vb
    return (vb.lbm, oracle.lbm) if tag == "ops2d" else (vb.lbm3d, oracle.lbm3d)


@pytest.mark.parametrize("tag", ["ops2d", "ops3d"])
def test_core_ops_vs_golden(golden, tag):
    g = golden[tag]
    m, _ = mods(tag)
    f, feq, rho, u, gg, om = T(g["f"]), T(g["feq"]), T(g["rho"]), T(g["u"]), T(g["g"]), float(g["omega"])
    assert_bitexact(N(m.streaming(f)), g["streaming"], "streaming")
    r, uu = m.get_macroscopic(f)
    assert_close(N(r), g["macro_rho"], what="rho"); assert_close(N(uu), g["macro_u"], what="u")
    r, uu = m.get_macroscopic(f[:, 1].contiguous())
    assert_close(N(r), g["macro_edge_rho"]); assert_close(N(uu), g["macro_edge_u"])
    assert_close(N(m.get_equilibrium(rho, u)), g["equilibrium"], what="feq")
    assert_close(N(m.collision_bgk(f, feq, om)), g["bgk"], what="bgk")
    assert_close(N(m.collision_kbc(f, feq, om)), g["kbc"], what="kbc")
    assert_close(N(m.collision_reg(f, feq, om)), g["reg"], what="reg")
    assert_close(N(m.collision_mrt(f, feq, g["mrt_op"])), g["mrt"], what="mrt")
    assert_close(N(m.collision_mrt(f, feq, m.get_mrt_collision_operator(om))), g["mrt"], what="mrt own operator")
    assert_close(N(m.get_guo_forcing_term(gg, u)), g["guo_term"], what="guo term")
    assert_close(N(m.forcing_edm(f, gg, u)), g["edm"], what="edm")
    assert_close(N(m.forcing_guo_bgk(f, gg, u, om)), g["guo_bgk"], what="guo bgk")
    assert_close(N(m.forcing_guo_mrt(f, gg, u, g["mrt_fop"])), g["guo_mrt"], what="guo mrt")
    assert_close(N(m.get_velocity_correction(gg, rho)), g["vel_corr"], what="velocity correction")
    assert_bitexact(N(m.obstacle_bounce_back(f, T(g["mask"]))), g["obstacle_bb"], "obstacle bounce-back")


@pytest.mark.parametrize("tag", ["ops2d", "ops3d"])
def test_boundaries_vs_golden(golden, tag):
    g = golden[tag]
    m, _ = mods(tag)
    comps = COMPS[tag]
    gcomps = [c.replace("u", "g", 1) for c in comps]
    f, f_pre = T(g["f"]), T(g["f_pre"])
    for loc in LOCS[tag]:
        uw_s = dict(zip(comps, g[f"{loc}_scalar_u"].tolist()))
        gw_s = dict(zip(gcomps, g[f"{loc}_scalar_g"].tolist()))
        uw_a = {c: T(g[f"{loc}_arr_{c}"]) for c in comps}
        rw_a = T(g[f"{loc}_arr_rho"])
        for kind in ("nee", "nebb", "equilibrium"):
            core = getattr(m, f"boundary_{kind}")
            assert_close(N(core(f, loc)), g[f"{kind}_{loc}_default"], what=f"{kind} {loc} default")
            assert_close(N(core(f, loc, rho_wall=1.02, **uw_s)), g[f"{kind}_{loc}_scalar"], what=f"{kind} {loc} scalar")
            assert_close(N(core(f, loc, rho_wall=rw_a, **uw_a)), g[f"{kind}_{loc}_array"], what=f"{kind} {loc} array")
            assert_close(N(getattr(m, f"boundary_velocity_{kind}")(f, loc, **uw_s)), g[f"velocity_{kind}_{loc}_scalar"],
                         what=f"velocity {kind} {loc}")
            assert_close(N(getattr(m, f"boundary_velocity_{kind}")(f, loc, **uw_a)), g[f"velocity_{kind}_{loc}_array"],
                         what=f"velocity {kind} {loc} array")
            assert_close(N(getattr(m, f"boundary_pressure_{kind}")(f, loc, rho_wall=0.98)), g[f"pressure_{kind}_{loc}_scalar"],
                         what=f"pressure {kind} {loc}")
            assert_close(N(getattr(m, f"boundary_force_corrected_{kind}")(f, loc, rho_wall=1.01, **uw_s, **gw_s)),
                         g[f"force_corrected_{kind}_{loc}_scalar"], what=f"force-corrected {kind} {loc}")
        assert_close(N(m.boundary_bounce_back(f_pre, f, loc)), g[f"bounce_back_{loc}_default"], what=f"bb {loc}")
        assert_close(N(m.boundary_bounce_back(f_pre, f, loc, **uw_s)), g[f"bounce_back_{loc}_scalar"], what=f"bb {loc} moving")
        assert_close(N(m.boundary_specular_reflection(f_pre, f, loc, **uw_s)), g[f"specular_{loc}_scalar"], what=f"specular {loc}")
        r, uu = m.boundary_characteristic(T(g["rho"]), T(g["u"]), loc)
        assert_close(N(r), g[f"cbc_{loc}_rho"], what=f"cbc rho {loc}"); assert_close(N(uu), g[f"cbc_{loc}_u"], what=f"cbc u {loc}")
    assert torch.equal(f, T(g["f"])), "boundary functions must not modify their input"
    with pytest.raises(KeyError):
        m.boundary_nee(f, "nowhere")
    with pytest.raises(ValueError):
        m.boundary_characteristic(T(g["rho"]), T(g["u"]), "nowhere")


def test_ib_vs_golden(golden):
    from vivsim_b200 import ib, ib3d
    g = golden["ib"]
    r = T(g["r"])
    assert_close(N(ib.kernel_peskin_3pt(r)), g["peskin3"]); assert_close(N(ib.kernel_peskin_4pt(r)), g["peskin4"])
    assert_close(N(ib.kernel_cosine_4pt(r)), g["cosine4"])
    assert_close(N(ib.kernel_hat_2pt(r)), oracle.ib.kernel_hat_2pt(g["r"]))
    mx, my, u = T(g["mx"]), T(g["my"]), T(g["u2"])
    ny = int(g["shape2"][1])
    for kname, kern in (("peskin4", ib.kernel_peskin_4pt), ("peskin3", ib.kernel_peskin_3pt), ("cosine4", ib.kernel_cosine_4pt)):
        w, idx = ib.get_ib_stencil(mx, my, ny, kernel=kern)
        assert_close(N(w), g[f"w2_{kname}"], what=f"weights {kname}")
        assert np.array_equal(N(idx), g[f"idx2_{kname}"]) and idx.dtype == torch.int32
    w, idx = ib.get_ib_stencil(mx, my, ny)
    assert_close(N(ib.interpolate(u, w, idx)), g["interp2"], what="interpolate")
    assert_close(N(ib.spread(T(g["vals2"]), u, w, idx)), g["spread2"], what="spread")
    for n_iter in (1, 5):
        gg, hh = ib.multi_direct_forcing(u, w, idx, T(g["tgt2"]), T(g["ds2_closed"]), n_iter=n_iter)
        assert_close(N(gg), g[f"mdf2_g_{n_iter}"], what="mdf g"); assert_close(N(hh), g[f"mdf2_h_{n_iter}"], what="mdf h")
    gg, hh = ib.multi_direct_forcing(u, w, idx, T(g["tgt2"]), 0.7, n_iter=3)
    assert_close(N(gg), g["mdf2_g_scalar_ds"]); assert_close(N(hh), g["mdf2_h_scalar_ds"])
    with pytest.raises(ValueError):
        ib.get_ib_stencil(mx, my, ny, kernel=lambda r: r)
    # 3-D
    verts, u3 = T(g["verts"]), T(g["u3"])
    shape = tuple(int(x) for x in g["shape3"])
    w, idx = ib3d.get_ib_stencil(verts, shape)
    assert_close(N(w), g["w3"]); assert np.array_equal(N(idx), g["idx3"])
    assert_close(N(ib3d.interpolate(u3, w, idx)), g["interp3"])
    assert_close(N(ib3d.spread(T(g["vals3"]), u3, w, idx)), g["spread3"])
    gg, hh = ib3d.multi_direct_forcing(u3, w, idx, T(g["tgt3"]), T(g["ds3"]), n_iter=3)
    assert_close(N(gg), g["mdf3_g"]); assert_close(N(hh), g["mdf3_h"])
    with pytest.raises(ValueError):
        ib3d.get_ib_stencil(verts, shape[:2])
    with pytest.raises(ValueError):
        ib3d.get_ib_stencil(verts[:, :2], shape)


@pytest.mark.parametrize("tag,shape", [("ops2d", (257, 130)), ("ops3d", (33, 18, 20)), ("ops2d", (1, 1)), ("ops3d", (2, 1, 3))])
def test_ops_vs_oracle_seeded(tag, shape):
    """Fresh seeded inputs at odd / degenerate sizes: CUDA vs the NumPy oracle."""
    m, o = mods(tag)
    rng = np.random.default_rng(5)
    dim = len(shape)
    rho = (1 + 0.05 * rng.standard_normal(shape)).astype(np.float32)
    u = (0.05 * rng.standard_normal((dim,) + shape)).astype(np.float32)
    feq = o.get_equilibrium(rho, u)
    f = (feq * (1 + 0.02 * rng.standard_normal(feq.shape))).astype(np.float32)
    gg = (1e-3 * rng.standard_normal((dim,) + shape)).astype(np.float32)
    assert_bitexact(N(m.streaming(T(f))), o.streaming(f), "streaming")
    r, uu = m.get_macroscopic(T(f)); ro, uo = o.get_macroscopic(f)
    assert_close(N(r), ro); assert_close(N(uu), uo)
    assert_close(N(m.get_equilibrium(T(rho), T(u))), feq)
    for name in ("bgk", "kbc", "reg"):
        assert_close(N(getattr(m, f"collision_{name}")(T(f), T(feq), 1.85)), getattr(o, f"collision_{name}")(f, feq, 1.85), what=name)
    op = o.get_mrt_collision_operator(1.85)
    assert_close(N(m.collision_mrt(T(f), T(feq), op)), o.collision_mrt(f, feq, op), what="mrt")
    assert_close(N(m.forcing_edm(T(f), T(gg), T(u))), o.forcing_edm(f, gg, u))
    mask = rng.random(shape) < 0.3
    assert_bitexact(N(m.obstacle_bounce_back(T(f), T(mask))), o.obstacle_bounce_back(f, mask), "mask")


def test_empty_inputs():
    from vivsim_b200 import lbm
    r, u = lbm.get_macroscopic(torch.zeros((9, 0), device="cuda"))
    assert r.shape == (0,) and u.shape == (2, 0)
    assert lbm.get_equilibrium(torch.zeros((0,), device="cuda"), torch.zeros((2, 0), device="cuda")).shape == (9, 0)


def test_ordered_spread_is_bit_exact_and_reproducible(golden):
    """vsb_ib_spread_ordered: contributions reach every cell in flattened (marker, stencil point) order,
    products and sums rounded separately -- bit-identical to the sequential scatter of the oracle (np.add.at,
    ib/stencil.py:104-110), and the same bits on every run; the atomic spread agrees to rounding only."""
    from vivsim_b200 import ib, ib3d
    import oracle.ib
    g = golden["ib"]
    # reference fixtures: 2-D and 3-D
    mx, my, u = T(g["mx"]), T(g["my"]), T(g["u2"])
    w, idx = ib.get_ib_stencil(mx, my, int(g["shape2"][1]))
    got = N(ib.spread(T(g["vals2"]), u, w, idx, ordered=True))
    assert_bitexact(got, oracle.ib.spread(g["vals2"], g["u2"], N(w), N(idx)), "ordered spread vs oracle (2-D fixture)")
    assert_close(got, g["spread2"], what="ordered spread vs reference fixture")
    verts, u3 = T(g["verts"]), T(g["u3"])
    w3, idx3 = ib3d.get_ib_stencil(verts, tuple(int(x) for x in g["shape3"]))
    got3 = N(ib3d.spread(T(g["vals3"]), u3, w3, idx3, ordered=True))
    assert_bitexact(got3, oracle.ib.spread(g["vals3"], g["u3"], N(w3), N(idx3)), "ordered spread vs oracle (3-D fixture)")
    # many markers on few cells (long runs of equal indices), negative and out-of-range indices, three components
    rng = np.random.default_rng(11)
    n_m, ns, ncell = 3000, 16, 40 * 25
    vals = rng.standard_normal((n_m, 3)).astype(np.float32)
    wts = rng.random((n_m, ns)).astype(np.float32)
    index = rng.integers(0, ncell, size=(n_m, ns)).astype(np.int32)
    grid = rng.standard_normal((3, 40, 25)).astype(np.float32)
    want = oracle.ib.spread(vals, grid, wts, index)
    runs = [N(ib.spread(T(vals), T(grid), T(wts), T(index), ordered=True)) for _ in range(3)]
    assert_bitexact(runs[0], want, "ordered spread vs oracle (dense duplicates)")
    assert_bitexact(runs[1], runs[0], "run to run"); assert_bitexact(runs[2], runs[0], "run to run")
    assert_close(N(ib.spread(T(vals), T(grid), T(wts), T(index))), want, what="atomic spread")
    index2 = index.copy()
    index2[::7, 0] = -1            # wraps once to the last cell
    index2[::11, 1] = ncell + 5    # dropped
    a = N(ib.spread(T(vals), T(grid), T(wts), T(index2), ordered=True))
    b = N(ib.spread(T(vals), T(grid), T(wts), T(index2)))
    assert_close(a, b, what="ordered vs atomic with wrapped / dropped indices")
    # empty marker set
    e = ib.spread(torch.zeros((0, 3), device="cuda"), T(grid), torch.zeros((0, ns), device="cuda"),
                  torch.zeros((0, ns), dtype=torch.int32, device="cuda"), ordered=True)
    assert_bitexact(N(e), grid, "no markers")
