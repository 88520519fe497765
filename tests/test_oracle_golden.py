"""CPU: the NumPy oracle against fixtures produced by the reference's own source
(tests/golden/make_golden.py).  This is what pins the oracle."""

import numpy as np
import pytest

from conftest import assert_bitexact, assert_close
from oracle import dyn, ib, ib3d, lbm, lbm3d, recipes
from oracle.lattice import D2Q9, D3Q19

MODS = {"ops2d": (lbm, D2Q9, ("left", "right", "top", "bottom"), ("ux_wall", "uy_wall")),
        "ops3d": (lbm3d, D3Q19, ("left", "right", "bottom", "top", "back", "front"),
                  ("ux_wall", "uy_wall", "uz_wall"))}


def test_lattice_tables(golden):
    g = golden["lattice"]
    for tag, lat in (("d2q9", D2Q9), ("d3q19", D3Q19)):
        assert np.array_equal(lat.c, g[f"{tag}_c"])
        assert np.array_equal(lat.w, g[f"{tag}_w"])
        assert np.array_equal(lat.opp, g[f"{tag}_opp"])
        assert np.array_equal(lat.opp[lat.opp], np.arange(lat.q))          # involution
        for loc, face in lat.faces.items():
            assert list(face.in_dirs) == list(g[f"{tag}_{loc}_in"]), loc
            assert list(face.out_dirs) == list(g[f"{tag}_{loc}_out"]), loc
            assert [face.sign, face.axis] == list(g[f"{tag}_{loc}_sign_axis"])
            if tag == "d2q9":
                assert list(face.tan_dirs) == list(g[f"{tag}_{loc}_tan"])
                assert sorted(face.pos_side_dirs) == sorted(g[f"{tag}_{loc}_pos"])
                assert sorted(face.neg_side_dirs) == sorted(g[f"{tag}_{loc}_neg"])
            else:
                assert list(face.zero_dirs) == list(g[f"{tag}_{loc}_zero"])
    assert np.array_equal(lbm.M, g["d2q9_M"])
    assert np.array_equal(lbm3d.M, g["d3q19_M"])
    for om in (0.8, 1.7):
        assert_close(lbm.get_mrt_collision_operator(om), g[f"d2q9_mrt_op_{om}"], what="mrt op 2d")
        assert_close(lbm.get_mrt_forcing_operator(om), g[f"d2q9_mrt_fop_{om}"], what="mrt fop 2d")
        assert_close(lbm3d.get_mrt_collision_operator(om), g[f"d3q19_mrt_op_{om}"], what="mrt op 3d")
        assert_close(lbm3d.get_mrt_forcing_operator(om), g[f"d3q19_mrt_fop_{om}"], what="mrt fop 3d")


def test_d3q19_projection_matches_closed_form(golden):
    P = golden["lattice"]["d3q19_P"]
    eye = np.eye(19, dtype=np.float32)
    mine = np.stack([lbm3d.core.second_order_projection(D3Q19, eye[:, r]) for r in range(19)], axis=1)
    assert_close(mine, P, what="second-order projection")


@pytest.mark.parametrize("tag", ["ops2d", "ops3d"])
def test_core_ops(golden, tag):
    g = golden[tag]
    m = MODS[tag][0]
    f, feq, rho, u, gg, om = g["f"], g["feq"], g["rho"], g["u"], g["g"], float(g["omega"])
    assert_bitexact(m.streaming(f), g["streaming"], "streaming")
    r, uu = m.get_macroscopic(f)
    assert_close(r, g["macro_rho"], what="rho"); assert_close(uu, g["macro_u"], what="u")
    r, uu = m.get_macroscopic(f[:, 1])
    assert_close(r, g["macro_edge_rho"], what="rho edge"); assert_close(uu, g["macro_edge_u"], what="u edge")
    assert_close(m.get_equilibrium(rho, u), g["equilibrium"], what="feq")
    assert_close(m.collision_bgk(f, feq, om), g["bgk"], what="bgk")
    assert_close(m.collision_kbc(f, feq, om), g["kbc"], what="kbc")
    assert_close(m.collision_reg(f, feq, om), g["reg"], what="reg")
    assert_close(m.get_mrt_collision_operator(om), g["mrt_op"], what="mrt op")
    assert_close(m.collision_mrt(f, feq, g["mrt_op"]), g["mrt"], what="mrt")
    assert_close(m.get_guo_forcing_term(gg, u), g["guo_term"], what="guo term")
    assert_close(m.forcing_edm(f, gg, u), g["edm"], what="edm")
    assert_close(m.forcing_guo_bgk(f, gg, u, om), g["guo_bgk"], what="guo bgk")
    assert_close(m.forcing_guo_mrt(f, gg, u, g["mrt_fop"]), g["guo_mrt"], what="guo mrt")
    assert_close(m.get_velocity_correction(gg, rho), g["vel_corr"], what="velocity correction")
    assert_bitexact(m.obstacle_bounce_back(f, g["mask"]), g["obstacle_bb"], "obstacle bounce-back")


@pytest.mark.parametrize("tag", ["ops2d", "ops3d"])
def test_boundaries(golden, tag):
    g = golden[tag]
    m, lat, locs, comps = MODS[tag]
    gcomps = [c.replace("u", "g", 1) for c in comps]
    f, f_pre = g["f"], g["f_pre"]
    for loc in locs:
        uw_s = dict(zip(comps, g[f"{loc}_scalar_u"].tolist()))
        gw_s = dict(zip(gcomps, g[f"{loc}_scalar_g"].tolist()))
        uw_a = {c: g[f"{loc}_arr_{c}"] for c in comps}
        rw_a = g[f"{loc}_arr_rho"]
        for kind in ("nee", "nebb", "equilibrium"):
            core = getattr(m, f"boundary_{kind}")
            assert_close(core(f, loc), g[f"{kind}_{loc}_default"], what=f"{kind} {loc} default")
            assert_close(core(f, loc, rho_wall=1.02, **uw_s), g[f"{kind}_{loc}_scalar"], what=f"{kind} {loc} scalar")
            assert_close(core(f, loc, rho_wall=rw_a, **uw_a), g[f"{kind}_{loc}_array"], what=f"{kind} {loc} array")
            assert_close(getattr(m, f"boundary_velocity_{kind}")(f, loc, **uw_s),
                         g[f"velocity_{kind}_{loc}_scalar"], what=f"velocity {kind} {loc}")
            assert_close(getattr(m, f"boundary_velocity_{kind}")(f, loc, **uw_a),
                         g[f"velocity_{kind}_{loc}_array"], what=f"velocity {kind} {loc} array")
            assert_close(getattr(m, f"boundary_pressure_{kind}")(f, loc, rho_wall=0.98),
                         g[f"pressure_{kind}_{loc}_scalar"], what=f"pressure {kind} {loc}")
            assert_close(getattr(m, f"boundary_force_corrected_{kind}")(f, loc, rho_wall=1.01, **uw_s, **gw_s),
                         g[f"force_corrected_{kind}_{loc}_scalar"], what=f"force-corrected {kind} {loc}")
        assert_close(m.boundary_bounce_back(f_pre, f, loc), g[f"bounce_back_{loc}_default"], what=f"bb {loc}")
        assert_close(m.boundary_bounce_back(f_pre, f, loc, **uw_s), g[f"bounce_back_{loc}_scalar"], what=f"bb {loc} moving")
        assert_close(m.boundary_specular_reflection(f_pre, f, loc, **uw_s), g[f"specular_{loc}_scalar"], what=f"specular {loc}")
        r, uu = m.boundary_characteristic(g["rho"], g["u"], loc)
        assert_close(r, g[f"cbc_{loc}_rho"], what=f"cbc rho {loc}"); assert_close(uu, g[f"cbc_{loc}_u"], what=f"cbc u {loc}")
    with pytest.raises(KeyError):
        m.boundary_nee(f, "nowhere")
    with pytest.raises(ValueError):
        m.boundary_characteristic(g["rho"], g["u"], "nowhere")


def test_ib_2d(golden):
    g = golden["ib"]
    assert_close(ib.kernel_peskin_3pt(g["r"]), g["peskin3"], what="peskin3")
    assert_close(ib.kernel_peskin_4pt(g["r"]), g["peskin4"], what="peskin4")
    assert_close(ib.kernel_cosine_4pt(g["r"]), g["cosine4"], what="cosine4")
    mx, my, u = g["mx"], g["my"], g["u2"]
    ny = int(g["shape2"][1])
    coords = np.stack([mx, my], axis=1)
    assert_close(ib.get_ds(coords), g["ds2_closed"], what="ds closed")
    assert_close(ib.get_ds(coords, closed=False), g["ds2_open"], what="ds open")
    assert_close(ib.get_area(coords), g["area2"], what="area")
    for kname, kern in recipes.KERNELS.items():
        if kname == "hat2":
            continue
        w, idx = ib.get_ib_stencil(mx, my, ny, kernel=kern)
        assert_close(w, g[f"w2_{kname}"], what=f"weights {kname}")
        assert np.array_equal(idx, g[f"idx2_{kname}"]) and idx.dtype == np.int32
    w, idx = ib.get_ib_stencil(mx, my, ny)
    assert_close(ib.interpolate(u, w, idx), g["interp2"], what="interpolate")
    assert_close(ib.spread(g["vals2"], u, w, idx), g["spread2"], what="spread")
    for n_iter in (1, 5):
        gg, hh = ib.multi_direct_forcing(u, w, idx, g["tgt2"], g["ds2_closed"], n_iter=n_iter)
        assert_close(gg, g[f"mdf2_g_{n_iter}"], what="mdf g"); assert_close(hh, g[f"mdf2_h_{n_iter}"], what="mdf h")
    gg, hh = ib.multi_direct_forcing(u, w, idx, g["tgt2"], 0.7, n_iter=3)
    assert_close(gg, g["mdf2_g_scalar_ds"], what="mdf g scalar ds"); assert_close(hh, g["mdf2_h_scalar_ds"], what="mdf h scalar ds")


def test_ib_3d(golden):
    g = golden["ib"]
    verts, faces, u3 = g["verts"], g["faces"], g["u3"]
    shape = tuple(int(x) for x in g["shape3"])
    assert_close(ib3d.get_triangle_areas(verts, faces), g["tri_areas"], what="tri areas")
    assert_close(ib3d.get_surface_area(verts, faces), g["surf_area"], what="surface area")
    assert_close(ib3d.get_volume(verts, faces), g["volume"], what="volume")
    assert_close(ib3d.get_ds(verts, faces), g["ds3"], what="ds3")
    w, idx = ib3d.get_ib_stencil(verts, shape)
    assert_close(w, g["w3"], what="w3"); assert np.array_equal(idx, g["idx3"])
    assert_close(ib3d.interpolate(u3, w, idx), g["interp3"], what="interp3")
    assert_close(ib3d.spread(g["vals3"], u3, w, idx), g["spread3"], what="spread3")
    gg, hh = ib3d.multi_direct_forcing(u3, w, idx, g["tgt3"], g["ds3"], n_iter=3)
    assert_close(gg, g["mdf3_g"], what="mdf3 g"); assert_close(hh, g["mdf3_h"], what="mdf3 h")
    with pytest.raises(ValueError):
        ib3d.get_ib_stencil(verts, shape[:2])
    with pytest.raises(ValueError):
        ib3d.get_ib_stencil(verts[:, :2], shape)


def test_dyn(golden):
    g = golden["dyn"]
    a, v, d, h = g["a"], g["v"], g["d"], g["h"]
    for got, key in zip(dyn.newmark_2dof(a, v, d, h, 31.4, 0.8, 0.05), ("a", "v", "d")):
        assert_close(got, g[f"nm_scalar_{key}"], what=f"newmark scalar {key}")
    for got, key in zip(dyn.newmark(a, v, d, h, g["m"], g["k"], g["c"]), ("a", "v", "d")):
        assert_close(got, g[f"nm_matrix_{key}"], what=f"newmark matrix {key}")
    x, y = dyn.get_markers_coords_2dof(g["x0"], g["y0"], d)
    assert_close(x, g["c2x"]); assert_close(y, g["c2y"])
    xm, ym = dyn.get_markers_coords_3dof(g["x0"], g["y0"], 6.0, 1.0, g["d3"])
    assert_close(xm, g["c3x"]); assert_close(ym, g["c3y"])
    assert_close(dyn.get_markers_velocity_3dof(xm, ym, 6.0, 1.0, g["d3"], g["v3"]), g["v3m"])
    assert_close(dyn.get_force_to_obj(g["hm"]), g["force"])
    assert_close(dyn.get_torque_to_obj(xm, ym, 6.0, 1.0, g["d3"], g["hm"]), g["torque"])
