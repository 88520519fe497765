"""CPU: the C + OpenMP restatement (bench CPU baseline / fast checker) against the NumPy oracle and the
reference-derived golden fixtures."""

import numpy as np
import pytest

import cases
from conftest import assert_close
from oracle import cport, recipes


def test_c2_and_kbc_cylinder_match_golden(golden):
    g = golden["recipes"]
    for recipe, key, hkey in (("c2", "c2_f10", "c2_h_last"), ("kbc_edm", "cyl_f25", None)):
        spec, f0, n, _ = cases.cylinder(g, recipe)
        r = cport.CRunner(spec, f0)
        assert_close(r.run(n), g[key], what=recipe)
        if hkey:
            assert_close(-r.marker_force, g[hkey], what="marker force")


def test_sphere_matches_golden(golden):
    g = golden["recipes"]
    spec, f0, n, key = cases.sphere(g)
    r = cport.CRunner(spec, f0)
    assert_close(r.run(n), g[key], what="sphere")
    assert_close(-r.marker_force, g["sphere_h_last"], what="marker force")


def test_viv_matches_golden(golden):
    g = golden["recipes"]
    spec, body, f0, (d, v, a), n = cases.viv(g)
    r = cport.CRunner(spec, f0, body=dict(body, d0=d, v0=v, a0=a))
    hist = []
    for _ in range(n):
        r.run(1)
        hist.append(np.concatenate(r.body_state()))
    assert_close(r.f, g["viv_f20"], what="viv f")
    assert_close(np.array(hist), g["viv_dvah"], rtol=1e-4, what="viv body history")


def test_thread_count_does_not_change_result():
    spec = recipes.cylinder2d_spec(nx=64, ny=48, n_marker=40, radius=6.0, u0=0.08, nu=0.02, n_iter=3)
    f0 = recipes.uniform_init(spec, noise=1e-3, seed=2)
    a = cport.CRunner(spec, f0).run(5, threads=1).copy()
    b = cport.CRunner(spec, f0).run(5, threads=4).copy()
    assert np.array_equal(a, b)
    f_np, _ = recipes.run(spec, f0, 5)
    assert_close(a, f_np, what="C port vs NumPy oracle")


def test_unsupported_recipe_raises(golden):
    spec, f0, _, _ = cases.cavity(golden["recipes"])
    with pytest.raises(NotImplementedError):
        cport.CRunner(spec, f0)


def test_mrt_guo_matches_golden_and_numpy(golden):
    """MRT collision + Guo-MRT source in the C port (C5's model): golden 3-D recipe and a 2-D IB case vs the NumPy oracle."""
    g = golden["recipes"]
    spec, f0, n, key = cases.mrt3(g)
    assert_close(cport.CRunner(spec, f0).run(n), g[key], what="mrt3")
    spec = recipes.cylinder2d_spec(nx=64, ny=48, n_marker=40, radius=6.0, u0=0.08, nu=0.02, n_iter=3, collision="mrt")
    f0 = recipes.uniform_init(spec, noise=1e-3, seed=3)
    f_np, h_np = recipes.run(spec, f0, 5)
    r = cport.CRunner(spec, f0)
    assert_close(r.run(5), f_np, what="2-D MRT + Guo-MRT + IB")
    assert_close(-r.marker_force, h_np, rtol=3e-5, what="marker force")


def test_moving_body_3d_clip_floor_window():
    """3-D elastically mounted cylinder (C5 recipe scaled down): 2-DOF Newmark body, window rule clip(floor())."""
    from vivsim_b200 import configs
    spec, body = configs.oscillating_cylinder_3d(nx=40, ny=28, nz=30, diameter=7.0, moving=True)
    f0 = recipes.uniform_init(spec, noise=1e-3, seed=4)
    # starts next to cell borders and crosses them within the run
    d, v, a = np.array([0.96, -0.03], np.float32), np.array([0.03, -0.02], np.float32), np.zeros(2, np.float32)
    r = cport.CRunner(spec, f0, body=dict(body, d0=d, v0=v, a0=a), follow=2)
    f = f0
    for k in range(4):
        f, d, v, a, h = recipes.viv_step(spec, body, f, d, v, a, follow=2)
        r.run(1)
        dc, vc, ac, hc = r.body_state()
        assert_close(dc, d, rtol=1e-5, what=f"d step {k}")
        assert_close(hc, h, rtol=1e-4, what=f"h step {k}")
    assert_close(r.f, f, what="3-D moving body f")


def test_fp64_yardstick_long_horizon_force_drift():
    """How far may two correct fp32 evaluations of the same recipe drift apart over 100 steps?  The C port in fp32, the
    NumPy oracle in fp32 and the C port in fp64 (same source, REF_REAL=double) on the C2 recipe at reduced size:
    populations stay within 1e-5, marker forces -- (U - u_m) 2 ds, a difference of nearly equal numbers -- sit at
    ~1e-5 of the largest force from the fp64 result, and the two fp32 oracles differ from EACH OTHER by about as much.
    This is the yardstick the GPU horizon tests use (conftest.assert_within_fp32_drift)."""
    from conftest import rel_err
    spec = recipes.cylinder2d_spec(nx=96, ny=64, n_marker=64, radius=7.5, u0=0.08, nu=0.02, n_iter=5)
    f0 = recipes.uniform_init(spec, noise=1e-3, seed=0)
    c32, c64 = cport.CRunner(spec, f0), cport.CRunner(spec, f0, dtype=np.float64)
    f32_, f64_ = c32.run(100).copy(), c64.run(100).copy()
    f_np, h_np = recipes.run(spec, f0, 100)
    assert rel_err(f32_, f64_) < 1e-5 and rel_err(f_np, f64_) < 1e-5
    d_c, d_np = rel_err(c32.marker_force, c64.marker_force), rel_err(-h_np, c64.marker_force)
    between = rel_err(-h_np, c32.marker_force)
    assert 1e-6 < d_c < 1e-4 and 1e-6 < d_np < 1e-4, (d_c, d_np)
    assert between <= 3 * max(d_c, d_np)
    # one step: fp32 and fp64 agree to fp32 rounding
    a, b = cport.CRunner(spec, f0), cport.CRunner(spec, f0, dtype=np.float64)
    a.run(1); b.run(1)
    assert rel_err(a.marker_force, b.marker_force) < 2e-6
