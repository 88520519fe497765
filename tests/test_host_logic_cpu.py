"""CPU: host-side logic of the hot path that needs no device -- the marker ordering / chunk cutting that feeds the tiled
MDF kernel, and the BASELINE workload generators."""

import numpy as np

from vivsim_b200 import configs
from vivsim_b200.stepper import TILE_CELLS, TILE_CHUNK, TILE_COLUMN, cut_marker_chunks


def check_chunks(markers, perm, offsets):
    m = markers.shape[0]
    assert sorted(perm.tolist()) == list(range(m)), "perm is not a permutation"
    assert offsets.dtype == np.int32 and offsets[0] == 0 and offsets[-1] == m
    sizes = np.diff(offsets)
    assert sizes.min() >= 1 and sizes.max() <= TILE_CHUNK
    srt = markers[perm]
    worst = 0
    for a, b in zip(offsets[:-1], offsets[1:]):
        c = srt[a:b]
        col = np.floor(c[:, :2] / TILE_COLUMN)
        assert (col == col[0]).all(), "a chunk spans more than one column"
        assert (np.diff(c[:, 2]) >= 0).all(), "markers of a chunk are not z-sorted"
        # bounding box of the 4-point stencils (floor(x) - 1 .. floor(x) + 2), + 1 cell per axis for a moving body
        lo = np.floor(c).min(axis=0) - 1
        hi = np.floor(c).max(axis=0) + 2
        ext = (hi - lo + 1) + 1
        worst = max(worst, int(np.prod(ext)))
    assert worst <= TILE_CELLS, f"a chunk's box needs {worst} cells"


def test_chunks_of_the_c5_cylinder_fit_the_tile():
    spec, _ = configs.oscillating_cylinder_3d()                       # BASELINE config 4: 695 570 markers
    markers = np.asarray(spec["ib"]["markers"], dtype=np.float32)
    assert markers.shape == (695570, 3)
    perm, offsets = cut_marker_chunks(markers)
    check_chunks(markers, perm, offsets)
    assert len(offsets) - 1 == 3496                                   # the grid size in profiles/r01_ncu_full_k_mdf_stage_tiled_c5_final.txt


def test_chunks_of_ragged_marker_sets():
    rng = np.random.default_rng(0)
    for n, box in ((1, 30.0), (7, 3.0), (481, 20.0), (5000, 40.0), (3000, 2.0)):
        markers = (rng.random((n, 3)) * box + 8.0).astype(np.float32)
        markers[:, 2] *= 4.0                                          # long in z: forces z cuts as well as 256-marker cuts
        perm, offsets = cut_marker_chunks(markers)
        check_chunks(markers, perm, offsets)
    # one column, one z: the 256-marker limit alone cuts it
    markers = np.tile(np.array([[9.5, 9.5, 20.25]], dtype=np.float32), (1000, 1))
    perm, offsets = cut_marker_chunks(markers)
    check_chunks(markers, perm, offsets)
    assert np.diff(offsets).tolist() == [256, 256, 256, 232]


def test_baseline_workload_generators():
    spec, body = configs.viv_cylinder_2d()                            # C2
    assert spec["shape"] == (1024, 1024) and spec["ib"]["markers"].shape == (512, 2) and spec["ib"]["n_iter"] == 5
    assert spec["collision"] == "bgk" and spec["forcing"] == "guo" and body is not None
    spec, _ = configs.sphere_3d()                                     # C3
    assert spec["shape"] == (256, 256, 256) and spec["ib"]["markers"].shape == (2562, 3) and spec["collision"] == "kbc"
    spec, _ = configs.viv_cylinder_2d_large()                         # C4
    assert spec["shape"] == (16384, 16384) and spec["collision"] == "kbc"
    for sp in (spec,):
        (ox, oy), (sx, sy) = sp["ib"]["window"]
        mk = np.asarray(sp["ib"]["markers"])
        # every 4-point stencil lies inside the window (the reference leaves out-of-range indices undefined)
        assert np.floor(mk[:, 0]).min() - 1 >= ox and np.floor(mk[:, 0]).max() + 2 < ox + sx
        assert np.floor(mk[:, 1]).min() - 1 >= oy and np.floor(mk[:, 1]).max() + 2 < oy + sy


def _moment_harness():
    """vivsim_b200/csrc/vsb_mrt_moment.cuh is host + device code: g++ compiles it into a small shared library."""
    import ctypes as C
    import os
    import subprocess
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "harness")
    out = os.path.join(here, "build", "libmrt_moment_harness.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    src = os.path.join(here, "mrt_moment_harness.cpp")
    hdr = os.path.join(here, "..", "..", "vivsim_b200", "csrc", "vsb_mrt_moment.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-I/usr/local/cuda/include", "-o", out, src], check=True)
    return C.CDLL(out)


def test_d3q19_mrt_in_moment_space_equals_the_reference_operator(golden):
    """The moment-space evaluation the fused D3Q19 MRT kernel uses (no matrix) against the dense product with the
    reference's own operators (lbm3d/collision/mrt.py:74-98): the rates are recovered from the matrix, A x agrees to
    fp32 rounding, the Guo operator must be I - A/2, and anything else is refused (-> matrix kernels)."""
    import ctypes as C
    lib = _moment_harness()
    fp = C.POINTER(C.c_float)
    P = lambda a: a.ctypes.data_as(fp)
    g = golden["lattice"]
    rng = np.random.default_rng(0)
    for om in ("0.8", "1.7"):
        A = np.ascontiguousarray(g[f"d3q19_mrt_op_{om}"], dtype=np.float32)
        B = np.ascontiguousarray(g[f"d3q19_mrt_fop_{om}"], dtype=np.float32)
        s = np.zeros(19, np.float32)
        assert lib.mm_make(P(A), P(B), P(s)) == 1
        want = [0, 0, 0, 0, 1.1] + [float(om)] * 5 + [1.2] * 6 + [1.4] * 3      # lbm3d/collision/mrt.py:50-72
        np.testing.assert_allclose(s, want, rtol=0, atol=2e-6)
        for _ in range(100):
            x = rng.standard_normal(19).astype(np.float32)
            y = np.zeros(19, np.float32)
            lib.mm_apply(P(s), P(x), P(y))
            ref = A.astype(np.float64) @ x.astype(np.float64)
            assert np.abs(y - ref).max() <= 2e-6 * np.abs(ref).max()
        bad = A.copy(); bad[3, 5] += 0.01
        assert lib.mm_make(P(bad), None, P(s)) == 0                    # not diagonal in the moment basis
        assert lib.mm_make(P(A), P(A), P(s)) == 0                      # source operator is not I - A/2
        cons = (A + 0.1 * np.eye(19, dtype=np.float32)).astype(np.float32)
        assert lib.mm_make(P(cons), None, P(s)) == 0                   # conserved moments relaxed


def test_reachable_window_cells_cover_every_stencil_of_a_following_window():
    """The cell list that lets the dense-body path skip most of the IB window: for any displacement of the body, with
    the window origin following it by floor() (examples/3d/oscillating_cylinder.py:241-243) or truncation
    (examples/2d/vortex_induced_vibration.py:104-105), every node of every marker's 4-point stencil is on the list."""
    from vivsim_b200.stepper import reachable_window_cells
    rng = np.random.default_rng(1)
    spec, _ = configs.oscillating_cylinder_3d(nx=128, ny=64, nz=64)
    markers = np.asarray(spec["ib"]["markers"], dtype=np.float32)
    origin, size = spec["ib"]["window"]
    cells = reachable_window_cells(markers, origin, size)
    assert cells.dtype == np.int32 and (np.diff(cells) > 0).all()
    assert cells.size < 0.6 * np.prod(size)                           # a shell, not the whole window
    listed = np.zeros(int(np.prod(size)), dtype=bool)
    listed[cells] = True
    for _ in range(20):
        d = rng.uniform(-3.0, 3.0, size=3).astype(np.float32)
        org = np.floor(np.asarray(origin, dtype=np.float32) + d).astype(np.int64)
        base = np.floor(markers + d - org.astype(np.float32)).astype(np.int64)
        for jx in range(-1, 3):
            for jy in range(-1, 3):
                for jz in range(-1, 3):
                    n = base + np.array([jx, jy, jz])
                    inside = ((n >= 0) & (n < np.asarray(size))).all(axis=1)
                    flat = (n[:, 0] * size[1] + n[:, 1]) * size[2] + n[:, 2]
                    assert listed[flat[inside]].all()
    # 2-D, truncation rule
    spec2, _ = configs.viv_cylinder_2d(nx=256, ny=128, n_marker=400, radius=20.0)
    m2 = np.asarray(spec2["ib"]["markers"], dtype=np.float32)
    (ox, oy), (sx, sy) = spec2["ib"]["window"]
    c2 = reachable_window_cells(m2, (ox, oy), (sx, sy))
    l2 = np.zeros(sx * sy, dtype=bool); l2[c2] = True
    for _ in range(20):
        d = rng.uniform(0.0, 2.0, size=2).astype(np.float32)
        org = (np.asarray((ox, oy), dtype=np.float32) + d).astype(np.int64)
        base = np.floor(m2 + d - org.astype(np.float32)).astype(np.int64)
        for jx in range(-1, 3):
            for jy in range(-1, 3):
                n = base + np.array([jx, jy])
                assert l2[n[:, 0] * sy + n[:, 1]].all()
