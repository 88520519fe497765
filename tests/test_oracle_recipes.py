"""CPU: composed oracle steps against the reference's own example recipes run to
the same horizon (fixtures from tests/golden/make_golden.py), plus invariants."""

import numpy as np
import pytest

import cases
from conftest import assert_close
from oracle import lbm, lbm3d, recipes
from oracle.core import F32

NAMES = ["cavity", "cavity_kbc_topfirst", "poiseuille_bgk_edm", "poiseuille_bgk_guo", "poiseuille_mrt_guo",
         "poiseuille_kbc_edm", "poiseuille_reg_edm", "cylinder_kbc_edm", "cylinder_c2", "text_mask",
         "sphere", "mrt3"]


@pytest.mark.parametrize("name", NAMES)
def test_recipe_matches_reference(golden, name):
    g = golden["recipes"]
    spec, f0, n, key = dict(cases.all_fluid_cases(g))[name]
    f, h = recipes.run(spec, f0, n)
    assert_close(f, g[key], what=name)
    if name == "cylinder_c2":
        assert_close(h, g["c2_h_last"], what="marker force")
    if name == "sphere":
        assert_close(h, g["sphere_h_last"], what="marker force")


def test_cylinder_force_history(golden):
    g = golden["recipes"]
    spec, f, n, _ = cases.cylinder(g, "kbc_edm")
    hs = []
    for _ in range(n):
        f, h = recipes.step(spec, f)
        hs.append(h.sum(axis=0))
    assert_close(np.array(hs), g["cyl_h"], what="force history")


def test_viv_moving_body(golden):
    g = golden["recipes"]
    spec, body, f, (d, v, a), n = cases.viv(g)
    hist = []
    for _ in range(n):
        f, d, v, a, h = recipes.viv_step(spec, body, f, d, v, a)
        hist.append(np.concatenate([d, v, a, h]))
    assert_close(f, g["viv_f20"], what="viv f")
    hist = np.array(hist); ref = g["viv_dvah"]
    for k, nm in enumerate(("d", "v", "a", "h")):
        assert_close(hist[:, 2 * k:2 * k + 2], ref[:, 2 * k:2 * k + 2], rtol=1e-4, what=f"viv {nm}")


def test_viv_rotating_body(golden):
    """3-DOF body (x, y, rotation; dyn.py:84-154 with the matrix-form Newmark) against the reference's own functions."""
    g = golden["rotation"]
    spec, body, f, (d, v, a), n = cases.viv_rotation(g)
    hist = []
    for _ in range(n):
        f, d, v, a, h = recipes.viv_step_3dof(spec, body, f, d, v, a)
        hist.append(np.concatenate([d, v, a, h]))
    assert_close(f, g["rot_f30"], what="rotation f")
    hist = np.array(hist); ref = g["rot_dvah"]
    for k, nm in enumerate(("d", "v", "a", "h")):
        assert_close(hist[:, 3 * k:3 * k + 3], ref[:, 3 * k:3 * k + 3], rtol=1e-4, what=f"rotation {nm}")


@pytest.mark.parametrize("mod,shape", [(lbm, (9, 7)), (lbm3d, (5, 4, 6))])
def test_invariants(mod, shape):
    rng = np.random.default_rng(0)
    dim = len(shape)
    rho = (1 + 0.05 * rng.standard_normal(shape)).astype(F32)
    u = (0.05 * rng.standard_normal((dim,) + shape)).astype(F32)
    feq = mod.get_equilibrium(rho, u)
    f = (feq * (1 + 0.02 * rng.standard_normal(feq.shape))).astype(F32)
    # streaming is a permutation of every population plane
    s = mod.streaming(f)
    for q in range(f.shape[0]):
        assert np.array_equal(np.sort(s[q].ravel()), np.sort(f[q].ravel()))
    # equilibrium reproduces its moments
    r2, u2 = mod.get_macroscopic(feq)
    assert_close(r2, rho); assert_close(u2, u, rtol=1e-4)
    # every collision conserves mass and momentum
    r0, u0 = mod.get_macroscopic(f)
    feq0 = mod.get_equilibrium(r0, u0)
    for out in (mod.collision_bgk(f, feq0, 1.6), mod.collision_kbc(f, feq0, 1.6), mod.collision_reg(f, feq0, 1.6),
                mod.collision_mrt(f, feq0, mod.get_mrt_collision_operator(1.6))):
        r1, u1 = mod.get_macroscopic(out)
        assert_close(r1, r0); assert_close(u1 * r1, u0 * r0, rtol=1e-4)
    # KBC at exact equilibrium leaves f unchanged (eps handling)
    assert_close(mod.collision_kbc(feq0, feq0, 1.9), feq0, rtol=1e-6)
    # Guo term: zeroth moment 0, first moment g
    g = (1e-3 * rng.standard_normal((dim,) + shape)).astype(F32)
    G = mod.get_guo_forcing_term(g, u)
    assert np.abs(G.sum(axis=0)).max() < 1e-8
