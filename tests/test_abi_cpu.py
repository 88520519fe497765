"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/vivsim_b200.h
declares; host-side validation fails loudly (no CPU fallback)."""

import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "vivsim_b200.h")


@pytest.fixture(scope="module")
def lib():
    from vivsim_b200 import _build, _lib
    _build.build()
    return _lib.lib()


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"^(?:int|int64_t|const char\*)\s+(vsb_\w+)\s*\(", src, flags=re.M)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 27
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    from vivsim_b200 import _lib
    assert sorted(_lib.EXPORTS) == names, "ctypes binding and header disagree"
    assert lib.vsb_abi_version() == 2


def test_struct_layouts_match_header(lib):
    from vivsim_b200 import _lib
    assert ctypes.sizeof(_lib.VsbBodyState) == 92
    assert ctypes.sizeof(_lib.VsbGrid) == 16
    assert ctypes.sizeof(_lib.VsbWallValue) == 16


def test_every_ctypes_struct_matches_the_header_field_by_field(tmp_path):
    """A C program compiled against include/vivsim_b200.h prints sizeof / offsetof of every structure and field the
    ctypes binding declares (a field the header does not have fails the compilation): both sides must agree."""
    import subprocess
    from vivsim_b200 import _lib
    structs = [v for k, v in vars(_lib).items()
               if k.startswith("Vsb") and isinstance(v, type) and issubclass(v, ctypes.Structure)]
    assert len(structs) >= 11
    src = ["#include <stdio.h>", "#include <stddef.h>", '#include "vivsim_b200.h"', "int main(void) {"]
    for st in structs:
        src.append(f'printf("{st.__name__} %zu\\n", sizeof({st.__name__}));')
        for name, *_ in st._fields_:
            src.append(f'printf("{st.__name__}.{name} %zu\\n", offsetof({st.__name__}, {name}));')
    src.append("return 0; }")
    (tmp_path / "layout.c").write_text("\n".join(src))
    exe = str(tmp_path / "layout")
    r = subprocess.run(["gcc", "-I", os.path.dirname(HEADER), str(tmp_path / "layout.c"), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    seen = 0
    for key, val in zip(out[::2], out[1::2]):
        if "." in key:
            st, field = key.split(".")
            expect = getattr(getattr(_lib, st), field).offset
        else:
            expect = ctypes.sizeof(getattr(_lib, key))
        assert int(val) == expect, f"{key}: header {val}, ctypes {expect}"
        seen += 1
    assert seen == sum(1 + len(st._fields_) for st in structs)


def test_validation_errors_without_gpu(lib):
    from vivsim_b200 import _lib
    g = _lib.VsbGrid(5, 4, 4, 1)
    rc = lib.vsb_streaming(ctypes.byref(g), ctypes.c_void_p(16), ctypes.c_void_p(32), None)
    assert rc == -1 and b"dim must be 2" in lib.vsb_last_error()
    rc = lib.vsb_collision(2, ctypes.c_int64(4), 1, ctypes.c_double(1.0), None, ctypes.c_void_p(16), ctypes.c_void_p(16),
                           ctypes.c_void_p(32), None)
    assert rc == -1 and b"MRT needs op_host" in lib.vsb_last_error()


def test_host_ode_runner_validates_before_touching_the_device(lib):
    """vsb_run_host_ode / _multi: argument checks come first and fail with a message (no GPU needed to see them)."""
    assert lib.vsb_run_host_ode(None, None, None, None, None, 4) == -1
    assert b"null argument" in lib.vsb_last_error()
    assert lib.vsb_run_host_ode_multi(0, None, None, None, None, None, 4) == -1
    assert b"n_domains must be 1..64" in lib.vsb_last_error()
    assert lib.vsb_run_host_ode_multi(65, None, None, None, None, None, 4) == -1
    assert lib.vsb_run_host_ode_multi(2, None, None, None, None, None, 4) == -1
    assert b"null argument" in lib.vsb_last_error()
    # a domain without a body state / mailbox is refused
    from vivsim_b200 import _lib
    a, m, bp, plan = _lib.VsbStepArgs(), _lib.VsbMdfArgs(), _lib.VsbBodyParams(), _lib.VsbHostPlan()
    pinned = (ctypes.c_float * 23)()
    arr = lambda t, x: (ctypes.POINTER(t) * 1)(ctypes.pointer(x))
    rc = lib.vsb_run_host_ode_multi(1, arr(_lib.VsbStepArgs, a), arr(_lib.VsbMdfArgs, m), arr(_lib.VsbBodyParams, bp),
                                    (ctypes.c_void_p * 1)(ctypes.addressof(pinned)), arr(_lib.VsbHostPlan, plan), 4)
    assert rc == -1 and b"needs a body state and a host mailbox" in lib.vsb_last_error()


def test_python_api_rejects_cpu_tensors(lib):
    from vivsim_b200 import VsbError, lbm
    with pytest.raises(VsbError):
        lbm.streaming(torch.zeros(9, 4, 4))
    with pytest.raises(TypeError):
        lbm.streaming(np.zeros((9, 4, 4), dtype=np.float32))


def test_mrt_operators_match_reference(lib, golden):
    from vivsim_b200 import lbm, lbm3d
    from vivsim_b200 import _api
    from conftest import assert_close
    g = golden["lattice"]
    assert np.array_equal(_api._basis(2), g["d2q9_M"])
    assert np.array_equal(_api._basis(3), g["d3q19_M"])
    for om in (0.8, 1.7):
        assert_close(lbm.get_mrt_collision_operator(om), g[f"d2q9_mrt_op_{om}"])
        assert_close(lbm.get_mrt_forcing_operator(om), g[f"d2q9_mrt_fop_{om}"])
        assert_close(lbm3d.get_mrt_collision_operator(om), g[f"d3q19_mrt_op_{om}"])
        assert_close(lbm3d.get_mrt_forcing_operator(om), g[f"d3q19_mrt_fop_{om}"])


def test_host_dyn_and_geometry(golden):
    from vivsim_b200 import dyn, ib, ib3d
    from conftest import assert_close
    g = golden["dyn"]
    for got, key in zip(dyn.newmark_2dof(g["a"], g["v"], g["d"], g["h"], 31.4, 0.8, 0.05), ("a", "v", "d")):
        assert_close(got, g[f"nm_scalar_{key}"])
    for got, key in zip(dyn.newmark(g["a"], g["v"], g["d"], g["h"], g["m"], g["k"], g["c"]), ("a", "v", "d")):
        assert_close(got, g[f"nm_matrix_{key}"])
    xm, ym = dyn.get_markers_coords_3dof(g["x0"], g["y0"], 6.0, 1.0, g["d3"])
    assert_close(xm, g["c3x"]); assert_close(ym, g["c3y"])
    assert_close(dyn.get_markers_velocity_3dof(xm, ym, 6.0, 1.0, g["d3"], g["v3"]), g["v3m"])
    assert_close(dyn.get_force_to_obj(g["hm"]), g["force"])
    assert_close(dyn.get_torque_to_obj(xm, ym, 6.0, 1.0, g["d3"], g["hm"]), g["torque"])
    gi = golden["ib"]
    coords = np.stack([gi["mx"], gi["my"]], axis=1)
    assert_close(ib.get_ds(coords), gi["ds2_closed"]); assert_close(ib.get_ds(coords, closed=False), gi["ds2_open"])
    assert_close(ib.get_area(coords), gi["area2"])
    assert_close(ib3d.get_ds(gi["verts"], gi["faces"]), gi["ds3"])
    assert_close(ib3d.get_volume(gi["verts"], gi["faces"]), gi["volume"])
    assert_close(ib3d.get_surface_area(gi["verts"], gi["faces"]), gi["surf_area"])


def test_reference_submodule_import_paths():
    """`from vivsim.ib.kernels import kernel_peskin_4pt`-style imports (used by the reference's own ib3d package and by
    its examples) resolve to the same objects as the package-level names."""
    import importlib
    import vivsim_b200
    paths = {"lbm": ["basic", "collision.kbc", "collision.mrt", "collision.reg", "forcing.guo", "forcing.edm",
                     "boundary.bb", "boundary.cbc", "boundary.eq", "boundary.nebb", "boundary.nee"],
             "ib": ["kernels", "stencil", "mdf", "geometry"], "ib3d": ["stencil", "geometry"]}
    paths["lbm3d"] = paths["lbm"]
    n = 0
    for pkg, subs in paths.items():
        top = getattr(vivsim_b200, pkg)
        for sub in subs:
            mod = importlib.import_module(f"vivsim_b200.{pkg}.{sub}")
            names = [k for k in vars(mod) if not k.startswith("_")]
            assert names, f"{mod.__name__} exports nothing"
            for k in names:
                assert getattr(mod, k) is getattr(top, k)
                n += 1
    assert n >= 60
    from vivsim_b200.lbm.boundary.nebb import boundary_velocity_nebb, boundary_force_corrected_nebb  # noqa: F401
    from vivsim_b200.ib.mdf import multi_direct_forcing  # noqa: F401


def test_xla_ffi_shim_type_checks_against_the_abi():
    """jaxlib (and with it xla/ffi/api/ffi.h) is absent from this image, so the XLA FFI handlers cannot be built; a
    minimal stand-in for that header (tests/harness/xla_mock) lets the compiler at least type-check every handler's use
    of the C ABI: struct fields, argument order and counts of the 21 entry points it forwards to."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "vivsim_b200", "csrc", "xla", "vivsim_b200_xla.cc")
    r = subprocess.run(["g++", "-fsyntax-only", "-std=c++17", "-Wall", "-I", os.path.join(root, "include"),
                        "-I", os.path.join(root, "tests", "harness", "xla_mock"), "-I", "/usr/local/cuda/include", src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = open(src).read()
    handlers = text.count("XLA_FFI_DEFINE_HANDLER_SYMBOL(")
    assert handlers >= 19
