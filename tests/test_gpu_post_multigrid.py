"""GPU: the post.py diagnostics and multigrid.py transfers through the C ABI against the
fixtures of the reference source and against the NumPy oracle on fresh inputs (odd sizes, minimum sizes, full size)."""

import numpy as np
import pytest
import torch

from conftest import assert_bitexact, assert_close, load_golden
import oracle.multigrid
import oracle.post

pytestmark = pytest.mark.gpu

FIELDS = ("velocity_magnitude", "velocity_gradient", "vorticity", "vorticity_magnitude", "divergence",
          "strain_rate", "strain_rate_magnitude", "kinetic_energy", "enstrophy", "q_criterion")
EXACT = ("velocity_gradient", "vorticity", "strain_rate")     # differences and exact halvings only


def T(x):
    return torch.as_tensor(np.ascontiguousarray(x), device="cuda")


def N(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def check_fields(post, u, ref_of, what):
    ut = T(u)
    gscale = float(np.abs(ref_of("velocity_gradient")).max())
    for name in FIELDS:
        got, ref = N(getattr(post, name)(ut)), ref_of(name)
        assert got.dtype == np.float32
        if name in EXACT:
            assert_bitexact(got, ref, f"{what} {name}")
        elif name in ("divergence", "q_criterion"):      # signed sums: relative to the gradient scale
            scale = gscale if name == "divergence" else gscale ** 2
            assert got.shape == ref.shape
            assert np.abs(got.astype(np.float64) - ref).max() <= 1e-5 * max(scale, 1e-30), f"{what} {name}"
        else:
            assert_close(got, ref, what=f"{what} {name}")


@pytest.mark.parametrize("tag", ["2d", "2d_thin", "3d", "3d_thin"])
def test_post_vs_golden(tag):
    from vivsim_b200 import post
    g = load_golden("post")
    u, rho = g[f"{tag}_u"], g[f"{tag}_rho"]
    check_fields(post, u, lambda name: g[f"{tag}_{name}"], tag)
    for name in ("mean_kinetic_energy", "mean_enstrophy"):
        got = getattr(post, name)(T(u))
        assert got.shape == () and got.dtype == torch.float32
        assert_close(N(got), g[f"{tag}_{name}"], what=name)
    assert_close(N(post.pressure(T(rho))), g[f"{tag}_pressure"], what="pressure")
    assert_close(N(post.pressure(T(rho), 0.25)), g[f"{tag}_pressure_cs2"], what="pressure cs2")
    assert_bitexact(N(post.calculate_curl(T(u))), g[f"{tag}_calculate_curl"])
    assert_bitexact(N(post.calculate_vorticity(T(u))), g[f"{tag}_calculate_vorticity"])
    assert_close(N(post.calculate_velocity_magnitude(T(u))), g[f"{tag}_calculate_velocity_magnitude"])
    assert_close(N(post.calculate_vorticity_dimensionless(T(u), 20.0, 0.05)), g[f"{tag}_vorticity_dimensionless"])


@pytest.mark.parametrize("shape", [(2, 2), (3, 257), (129, 5), (300, 301), (2, 2, 2), (5, 3, 33), (33, 34, 35)])
def test_post_vs_oracle(shape):
    from vivsim_b200 import post
    rng = np.random.default_rng(sum(shape))
    u = (0.1 * rng.standard_normal((len(shape),) + shape)).astype(np.float32)
    check_fields(post, u, lambda name: getattr(oracle.post, name)(u), str(shape))
    assert_close(N(post.mean_kinetic_energy(T(u))), oracle.post.mean_kinetic_energy(u), what="mean ke")
    assert_close(N(post.mean_enstrophy(T(u))), oracle.post.mean_enstrophy(u), what="mean enstrophy")


def test_post_errors_and_full_size():
    from vivsim_b200 import post
    from vivsim_b200._lib import VsbError
    with pytest.raises(ValueError):
        post.vorticity(torch.zeros((2, 1, 8), device="cuda"))          # jnp.gradient needs 2 cells per axis
    with pytest.raises(ValueError):
        post.vorticity(torch.zeros((3, 8, 8), device="cuda"))          # leading axis must equal the dimension
    with pytest.raises(VsbError):
        post.vorticity(torch.zeros((2, 8, 8)))                         # no CPU path
    assert N(post.kinetic_energy(torch.ones((2, 1, 8), device="cuda"))).shape == (1, 8)   # no gradient: any size
    # BASELINE C2 size, properties that need no oracle: a rigid rotation has vorticity 2 Omega, zero strain and
    # zero divergence everywhere including the one-sided edges; Q = Omega^2
    n = 1024
    x = torch.arange(n, device="cuda", dtype=torch.float32)
    rot = torch.stack([(-0.25 * x)[None, :].expand(n, n), (0.25 * x)[:, None].expand(n, n)]).contiguous()
    assert torch.equal(post.vorticity(rot), torch.full((n, n), 0.5, device="cuda"))
    assert torch.equal(post.divergence(rot), torch.zeros((n, n), device="cuda"))
    assert torch.equal(post.strain_rate_magnitude(rot), torch.zeros((n, n), device="cuda"))
    assert torch.equal(post.q_criterion(rot), torch.full((n, n), 0.0625, device="cuda"))
    assert abs(float(post.mean_enstrophy(rot)) - 0.125) < 1e-7


def test_post_of_stepper_state():
    """Diagnostics of a running simulation: rho, u of the stepper's state -> vorticity, against the oracle."""
    from oracle import recipes
    from vivsim_b200 import Stepper, lbm, post
    spec = recipes.cylinder2d_spec(nx=96, ny=64, n_marker=64, radius=7.5, u0=0.08, nu=0.02, n_iter=3)
    f0 = recipes.uniform_init(spec, noise=1e-3, seed=0)
    st = Stepper(spec).set_f(f0)
    st.step(20)
    f = st.get_f()
    rho, u = lbm.get_macroscopic(f)
    f_ref, _ = recipes.run(spec, f0, 20)
    _, u_ref = oracle.lbm.get_macroscopic(f_ref)
    w, w_ref = N(post.vorticity(u)), oracle.post.vorticity(u_ref)
    assert np.abs(w - w_ref).max() <= 1e-5 * np.abs(u_ref).max()      # gradient of a field that agrees to 1e-5
    assert np.abs(w_ref).max() > 1e-3                                   # the wake exists


@pytest.mark.parametrize("tag,dirs", [("lr", ("left", "right")), ("ud", ("up", "down")), ("lr_min", ("left", "right")),
                                      ("ud_min", ("up", "down"))])
def test_multigrid_vs_golden(tag, dirs):
    from vivsim_b200 import multigrid
    g = load_golden("multigrid")
    ff, fc = g[f"{tag}_fine"], g[f"{tag}_coarse"]
    for d in dirs:
        ffc, fcc = T(ff), T(fc)
        assert_bitexact(N(multigrid.fine_to_coarse(ffc, fcc, d)), g[f"{tag}_f2c_{d}"], f"f2c {d}")
        assert_bitexact(N(multigrid.coarse_to_fine(fcc, ffc, d)), g[f"{tag}_c2f_{d}"], f"c2f {d}")
        assert_bitexact(N(ffc), ff, "inputs untouched"); assert_bitexact(N(fcc), fc, "inputs untouched")
    assert_bitexact(N(multigrid.fine_to_coarse(T(ff), T(fc), "top")), fc, "unknown dir: unchanged")


@pytest.mark.parametrize("nc", [(1, 1), (3, 129), (257, 2), (512, 512)])
def test_multigrid_vs_oracle(nc):
    from vivsim_b200 import multigrid
    rng = np.random.default_rng(nc[0] * 1000 + nc[1])
    fc = rng.standard_normal((9,) + nc).astype(np.float32)
    for d in ("left", "right", "up", "down"):
        fine_shape = (max(2, nc[0] + 1), 2 * nc[1]) if d in ("left", "right") else (2 * nc[0], max(2, nc[1] + 3))
        ff = rng.standard_normal((9,) + fine_shape).astype(np.float32)
        assert_bitexact(N(multigrid.fine_to_coarse(T(ff), T(fc), d)), oracle.multigrid.fine_to_coarse(ff, fc, d), d)
        assert_bitexact(N(multigrid.coarse_to_fine(T(fc), T(ff), d)), oracle.multigrid.coarse_to_fine(fc, ff, d), d)


def test_multigrid_host_helpers_and_errors():
    from vivsim_b200 import multigrid
    g = load_golden("multigrid")
    for nu, lv, om in g["omega"]:
        assert multigrid.get_omega(nu, int(lv)) == pytest.approx(om, rel=1e-12)
    for lv, (ix, iy) in zip((-1, 0, 1, 2), g["coord"]):
        assert multigrid.coord_to_indices(13.5, 7.25, 4, 2, lv) == (ix, iy)
    for i in range(3):
        w, h, lv, bx, by = (int(v) for v in g[f"init{i}_args"])
        f, rho, u = multigrid.init_grid(w, h, lv, bx, by)
        assert list(f.shape) + list(rho.shape) + list(u.shape) == list(g[f"init{i}_shapes"])
        assert [float(f.sum()), float(rho.mean()), float(u.sum())] == list(g[f"init{i}_sums"])
        assert f.is_cuda and f.dtype == torch.float32
    with pytest.raises(ValueError):
        multigrid.fine_to_coarse(torch.zeros((9, 4, 7), device="cuda"), torch.zeros((9, 4, 3), device="cuda"), "left")
    with pytest.raises(ValueError):
        multigrid.coarse_to_fine(torch.zeros((9, 3, 4), device="cuda"), torch.zeros((9, 5, 4), device="cuda"), "up")
