"""CPU: the NumPy oracle of post.py / multigrid.py against fixtures produced by the
reference's own source (tests/golden/make_golden.py post multigrid)."""

import numpy as np
import pytest

from conftest import assert_bitexact, assert_close, load_golden
from oracle import multigrid, post

POST_TAGS = ("2d", "2d_thin", "3d", "3d_thin")
POST_FIELDS = ("velocity_magnitude", "velocity_gradient", "vorticity", "vorticity_magnitude", "divergence",
               "strain_rate", "strain_rate_magnitude", "kinetic_energy", "enstrophy", "q_criterion")


@pytest.fixture(scope="module")
def gpost():
    return load_golden("post")


@pytest.fixture(scope="module")
def gmg():
    return load_golden("multigrid")


@pytest.mark.parametrize("tag", POST_TAGS)
def test_post_fields(gpost, tag):
    u = gpost[f"{tag}_u"]
    scale_g = np.abs(gpost[f"{tag}_velocity_gradient"]).max()
    for name in POST_FIELDS:
        got, ref = getattr(post, name)(u), gpost[f"{tag}_{name}"]
        assert got.dtype == np.float32 and ref.dtype == np.float32, name
        if name in ("velocity_gradient", "vorticity", "strain_rate"):
            assert_bitexact(got, ref, f"{tag} {name}")     # differences and exact halvings only
        elif name in ("divergence", "q_criterion"):
            # sums of signed gradient products: 1e-5 of the gradient scale (cancellation makes the field small)
            ref_scale = scale_g if name == "divergence" else scale_g ** 2
            assert np.abs(got.astype(np.float64) - ref).max() <= 1e-5 * ref_scale, name
        else:
            assert_close(got, ref, what=f"{tag} {name}")
    assert_close(post.mean_kinetic_energy(u), gpost[f"{tag}_mean_kinetic_energy"], what="mean ke")
    assert_close(post.mean_enstrophy(u), gpost[f"{tag}_mean_enstrophy"], what="mean enstrophy")
    assert_close(post.pressure(gpost[f"{tag}_rho"]), gpost[f"{tag}_pressure"], what="pressure")
    assert_close(post.pressure(gpost[f"{tag}_rho"], 0.25), gpost[f"{tag}_pressure_cs2"], what="pressure cs2")
    # the deprecated aliases are the same functions (post.py:180-211)
    assert_bitexact(gpost[f"{tag}_calculate_curl"], gpost[f"{tag}_vorticity"])
    assert_bitexact(gpost[f"{tag}_calculate_vorticity"], gpost[f"{tag}_vorticity"])
    assert_bitexact(gpost[f"{tag}_calculate_velocity_magnitude"], gpost[f"{tag}_velocity_magnitude"])
    assert_close(post.vorticity(u) * np.float32(20.0) / np.float32(0.05), gpost[f"{tag}_vorticity_dimensionless"],
                 what="dimensionless vorticity")


def test_post_identities():
    """Properties that hold for any field: Q = 0.5 (|W|^2 - |S|^2), a linear field has a constant gradient (also at the
    one-sided edges), a rigid rotation has vorticity 2 Omega and zero strain."""
    rng = np.random.default_rng(0)
    for shape in ((9, 8), (6, 5, 4)):
        d = len(shape)
        u = rng.standard_normal((d,) + shape).astype(np.float32)
        G = post.velocity_gradient(u).astype(np.float64)
        S = 0.5 * (G + np.swapaxes(G, 0, 1))
        W = 0.5 * (G - np.swapaxes(G, 0, 1))
        q = 0.5 * ((W * W).sum((0, 1)) - (S * S).sum((0, 1)))
        assert np.abs(post.q_criterion(u) - q).max() < 1e-5 * np.abs(G).max() ** 2
        grids = np.meshgrid(*[np.arange(n, dtype=np.float32) for n in shape], indexing="ij")
        A = rng.integers(-3, 4, size=(d, d)).astype(np.float32)
        lin = np.stack([sum(A[i, j] * grids[j] for j in range(d)) for i in range(d)])
        Gl = post.velocity_gradient(lin)
        for i in range(d):
            for j in range(d):
                assert np.array_equal(Gl[i, j], np.full(shape, A[i, j], dtype=np.float32))
    x, y = np.meshgrid(np.arange(7, dtype=np.float32), np.arange(6, dtype=np.float32), indexing="ij")
    rot = np.stack([-0.5 * y, 0.5 * x])
    assert np.array_equal(post.vorticity(rot), np.ones((7, 6), dtype=np.float32))
    assert np.array_equal(post.strain_rate_magnitude(rot), np.zeros((7, 6), dtype=np.float32))
    with pytest.raises(ValueError):
        post.vorticity(np.zeros((2, 1, 5), dtype=np.float32))


@pytest.mark.parametrize("tag,dirs", [("lr", ("left", "right")), ("ud", ("up", "down")), ("lr_min", ("left", "right")),
                                      ("ud_min", ("up", "down"))])
def test_multigrid_transfers(gmg, tag, dirs):
    ff, fc = gmg[f"{tag}_fine"], gmg[f"{tag}_coarse"]
    for d in dirs:
        assert_bitexact(multigrid.fine_to_coarse(ff, fc, d), gmg[f"{tag}_f2c_{d}"], f"f2c {d}")
        assert_bitexact(multigrid.coarse_to_fine(fc, ff, d), gmg[f"{tag}_c2f_{d}"], f"c2f {d}")
        # only the three populations crossing that edge change, and only on the receiving line
        changed = np.nonzero((multigrid.fine_to_coarse(ff, fc, d) != fc).any(axis=(1, 2)))[0]
        assert set(changed) <= set(multigrid.DIRS[d])
    assert_bitexact(multigrid.fine_to_coarse(ff, fc, "top"), gmg[f"{tag}_f2c_other"], "unknown dir")
    assert_bitexact(gmg[f"{tag}_f2c_other"], fc, "unknown dir leaves f_coarse unchanged")


def test_multigrid_roundtrip_and_host_helpers(gmg):
    """coarse -> fine -> coarse returns the coarse line (mean of four copies), and the host helpers match."""
    rng = np.random.default_rng(3)
    fc = rng.standard_normal((9, 5, 6)).astype(np.float32)
    # left transfer: coarse column 0 -> fine last column; pad a second fine layer with the same values
    ff = np.zeros((9, 4, 12), dtype=np.float32)
    ff = multigrid.coarse_to_fine(fc, ff, "left")
    ff[:, 0], ff[:, 1] = ff[:, -1], ff[:, -1]
    back = multigrid.fine_to_coarse(ff, np.zeros_like(fc), "left")
    q = multigrid.DIRS["left"]
    assert np.array_equal(back[q, -1], fc[q, 0])
    for nu, lv, om in gmg["omega"]:
        assert multigrid.get_omega(nu, int(lv)) == pytest.approx(om, rel=1e-12)
    assert multigrid.get_omega(0.05, 0) == pytest.approx(1 / (3 * 0.05 + 0.5))
    for lv, (ix, iy) in zip((-1, 0, 1, 2), gmg["coord"]):
        assert multigrid.coord_to_indices(13.5, 7.25, 4, 2, lv) == (ix, iy)
