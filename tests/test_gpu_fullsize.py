"""GPU: the fused stepper against the oracle AT THE BENCHMARKED SIZES.

The reduced-size recipe tests never reach the paths that only large grids take: wall blocks appended to many waves of
bulk blocks (edges = 2), the L2 prefetch, the 3-D force-window band with its own row decomposition, the tiled MDF kernel
with thousands of chunks.  Here every BASELINE configuration runs at its stated size (or the largest one the C oracle
finishes in seconds) on the same inputs through both sides:

    C2  D2Q9 BGK + Guo, 1024^2, 512 markers, moving body          20 steps   (examples/2d/vortex_induced_vibration.py:96-148)
    C3  D3Q19 KBC + EDM, 256^3, 2562-marker sphere                 3 steps   (examples/3d/flow_past_sphere.py:139-171)
    C4  D2Q9 KBC + EDM VIV recipe at 4096^2                        10 steps
    C5  D3Q19 MRT + Guo-MRT, 256 x 128 x 128, dense cylinder       4 steps   (examples/3d/oscillating_cylinder.py:229-282)

The oracle is oracle.cport (C + OpenMP restatement, itself checked against the golden fixtures and the NumPy oracle in
the CPU suite).  Tolerance: 1e-5 of the field scale for populations (north_star); 3e-5 for marker forces after several
steps (differences of nearly equal numbers; see conftest.assert_within_fp32_drift and the 100-step tests)."""

import numpy as np
import pytest
import torch

from conftest import assert_close
from oracle import cport

pytestmark = pytest.mark.gpu


def N(x):
    return x.detach().cpu().numpy()


def _state(spec, seed):
    """Perturbed equilibrium, built by the product's own get_equilibrium and handed to both sides."""
    from vivsim_b200 import configs
    return configs.uniform_state(spec, noise=1e-3, seed=seed)


def _compare_body(st, runner, what, rtol=1e-4):
    d, v, a, h = st.body_state()
    dr, vr, ar, hr = runner.body_state()
    assert_close(d, dr, rtol=rtol, what=f"{what}: displacement")
    assert_close(v, vr, rtol=rtol, what=f"{what}: velocity")
    assert_close(h, hr, rtol=rtol, what=f"{what}: hydrodynamic force")


@pytest.mark.parametrize("mode", ["device", "host", "graph"])
def test_c2_full_size_moving_body(mode):
    from vivsim_b200 import Stepper, configs
    spec, body = configs.viv_cylinder_2d()                      # 1024^2, 512 markers, MDF(5) + Guo, NEBB / equilibrium
    assert tuple(spec["shape"]) == (1024, 1024)
    f0 = _state(spec, 0)
    n = 20
    ref = cport.CRunner(spec, N(f0), body=body, follow=1)
    f_ref = ref.run(n)
    kw = dict(dyn_mode="host") if mode == "host" else dict(dyn_mode="device", use_graph=(mode == "graph"))
    st = Stepper(spec, body=dict(body), follow=1, **kw).set_f(f0)
    assert st.edge_fused and st.overlap
    st.step(n)
    assert_close(N(st.get_f()), f_ref, what=f"C2 1024^2 x {n} steps ({mode})")
    assert_close(N(st.marker_force), ref.marker_force, rtol=3e-5, what="C2 marker forces")
    _compare_body(st, ref, f"C2 ({mode})")


def test_c3_full_size_sphere():
    from vivsim_b200 import Stepper, configs
    spec, _ = configs.sphere_3d()                               # 256^3, 2562 markers, KBC, MDF(3) + EDM
    assert tuple(spec["shape"]) == (256, 256, 256)
    f0 = _state(spec, 1)
    n = 3
    ref = cport.CRunner(spec, N(f0))
    f_ref = ref.run(n)
    st = Stepper(spec).set_f(f0)
    st.step(n)
    got = st.get_f()
    # compare on the device to keep the host copy count down (2.5 GB per state)
    r = torch.as_tensor(f_ref, device="cuda")
    err = float((got - r).abs().max() / r.abs().max())
    assert err <= 1e-5, f"C3 256^3 x {n} steps: max |diff| / max |ref| = {err:.3e}"
    assert_close(N(st.marker_force), ref.marker_force, rtol=3e-5, what="C3 marker forces")


def test_c4_recipe_4096():
    from vivsim_b200 import Stepper, configs
    spec, body = configs.viv_cylinder_2d_large(n=4096)          # KBC + EDM, D = n/20, 4D markers, MDF(5), moving body
    f0 = _state(spec, 2)
    n = 10
    ref = cport.CRunner(spec, N(f0), body=body, follow=1)
    f_ref = ref.run(n)
    st = Stepper(spec, body=dict(body), dyn_mode="device", follow=1).set_f(f0)
    st.step(n)
    r = torch.as_tensor(f_ref, device="cuda")
    err = float((st.get_f() - r).abs().max() / r.abs().max())
    assert err <= 1e-5, f"C4 4096^2 x {n} steps: max |diff| / max |ref| = {err:.3e}"
    assert_close(N(st.marker_force), ref.marker_force, rtol=3e-5, what="C4 marker forces")
    _compare_body(st, ref, "C4")


@pytest.mark.parametrize("moving", [False, True])
def test_c5_recipe_256x128x128_tiled_mdf(moving):
    from vivsim_b200 import Stepper, configs
    spec, body = configs.oscillating_cylinder_3d(nx=256, ny=128, nz=128, moving=moving)   # MRT + Guo-MRT, MDF(3)
    M = spec["ib"]["markers"].shape[0]
    assert M > 20000
    f0 = _state(spec, 3)
    n = 4
    ref = cport.CRunner(spec, N(f0), body=body, follow=2)
    f_ref = ref.run(n)
    st = (Stepper(spec, body=dict(body), dyn_mode="device", follow=2) if moving else Stepper(spec)).set_f(f0)
    assert st._use_uwin and st._perm is not None and st._mdf.n_chunks > 100      # the tiled kernel with cut chunks
    st.step(n)
    r = torch.as_tensor(f_ref, device="cuda")
    err = float((st.get_f() - r).abs().max() / r.abs().max())
    assert err <= 1e-5, f"C5 256x128x128 x {n} steps: max |diff| / max |ref| = {err:.3e}"
    assert_close(N(st.marker_force), ref.marker_force, rtol=3e-5, what="C5 marker forces (caller's order)")
    if moving:
        _compare_body(st, ref, "C5")
