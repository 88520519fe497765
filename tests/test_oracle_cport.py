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
