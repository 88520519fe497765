"""ctypes wrapper of oracle/c/iblbm_ref.c (C + OpenMP restatement of the composed step).
TEST INFRASTRUCTURE ONLY: a fast checker for long / large runs and bench.py's CPU baseline.

Covers the recipes of BASELINE configs 2-5: collision bgk|mrt|kbc|reg, forcing none|edm|guo (Guo-MRT with MRT),
uniform g, one immersed body (Peskin 4-point, MDF, optional 2-DOF Newmark in 2-D and 3-D, window rule trunc or
clip(floor)), inlet NEBB / outlet equilibrium on the x faces or fully periodic.  Anything else raises."""

import ctypes as C
import os

import numpy as np

from .c import build as _build

_COLL = {"bgk": 0, "mrt": 1, "kbc": 2, "reg": 3}
_FORCE = {None: 0, "edm": 1, "guo": 2}


def _refspec(real):
    """ctypes mirror of RefSpec for REF_REAL = float | double."""
    class RefSpec(C.Structure):
        _fields_ = [("dim", C.c_int), ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("collision", C.c_int),
                    ("forcing", C.c_int), ("omega", C.c_double), ("g0", real * 3), ("n_markers", C.c_int),
                    ("n_iter", C.c_int), ("markers0", C.c_void_p), ("ds", C.c_void_p), ("worg0", real * 3),
                    ("wsz", C.c_int * 3), ("moving", C.c_int), ("body_m", C.c_double), ("body_k", C.c_double),
                    ("body_c", C.c_double), ("body_added", C.c_double), ("d", real * 3), ("v", real * 3),
                    ("a", real * 3), ("h", real * 3), ("inlet_outlet", C.c_int), ("u0", real),
                    ("marker_force", C.c_void_p), ("mrt_op", C.c_void_p), ("mrt_fop", C.c_void_p), ("follow", C.c_int)]
    return RefSpec


RefSpec = _refspec(C.c_float)
RefSpec64 = _refspec(C.c_double)

_lib = None
_lib64 = None


def lib(dtype=np.float32):
    global _lib, _lib64
    if np.dtype(dtype) == np.float64:
        if _lib64 is None:
            _lib64 = C.CDLL(_build.build64())
            _lib64.ref_num_threads.restype = C.c_int
        return _lib64
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.ref_num_threads.restype = C.c_int
    return _lib


def num_threads():
    return int(lib().ref_num_threads())


class CRunner:
    """Holds the buffers of one simulation so repeated ``run`` calls can be timed without allocation."""

    def __init__(self, spec, f0, body=None, follow=1, dtype=np.float32):
        """dtype=np.float64 runs the same source compiled with REF_REAL=double (the fp64 yardstick)."""
        dim = spec["dim"]
        shape = tuple(spec["shape"])
        self.dtype = dt = np.dtype(dtype)
        s = RefSpec64() if dt == np.float64 else RefSpec()
        s.dim, s.nx, s.ny, s.nz = dim, shape[0], shape[1], (shape[2] if dim == 3 else 1)
        if spec["collision"] not in _COLL:
            raise NotImplementedError(f"C port: collision {spec['collision']!r}")
        s.collision, s.forcing, s.omega = _COLL[spec["collision"]], _FORCE[spec.get("forcing")], float(spec["omega"])
        s.follow = int(follow)
        if spec["collision"] == "mrt":      # the operators the reference builds on the host (lbm/collision/mrt.py:47-63)
            from . import lbm, lbm3d
            mod = lbm if dim == 2 else lbm3d
            op = spec.get("mrt_op")
            self._op = np.ascontiguousarray(op if op is not None else mod.get_mrt_collision_operator(s.omega), dtype=dt)
            s.mrt_op = self._op.ctypes.data
            if spec.get("forcing") == "guo":
                fop = spec.get("mrt_fop")
                self._fop = np.ascontiguousarray(fop if fop is not None else mod.get_mrt_forcing_operator(s.omega),
                                                 dtype=dt)
                s.mrt_fop = self._fop.ctypes.data
        g = spec.get("g")
        if g is not None:
            g = np.asarray(g, dtype=np.float32)
            if g.ndim != 1:
                raise NotImplementedError("C port: only a uniform body force")
            for d in range(dim):
                s.g0[d] = float(g[d])
        self.marker_force = None
        ib = spec.get("ib")
        if ib is not None:
            if ib.get("kernel", "peskin4") != "peskin4" or ib.get("u_target") is not None:
                raise NotImplementedError("C port: Peskin 4-point kernel, zero / body target velocity only")
            self._markers = np.ascontiguousarray(np.asarray(ib["markers"], dtype=np.float32), dtype=dt)
            m = self._markers.shape[0]
            self._ds = np.ascontiguousarray(np.broadcast_to(np.asarray(ib["ds"], dtype=np.float32), (m,)), dtype=dt)
            self.marker_force = np.zeros((m, dim), dtype=dt)
            s.n_markers, s.n_iter = m, int(ib.get("n_iter", 5))
            s.markers0, s.ds = self._markers.ctypes.data, self._ds.ctypes.data
            s.marker_force = self.marker_force.ctypes.data
            for d in range(dim):
                s.worg0[d] = float(ib["window"][0][d]); s.wsz[d] = int(ib["window"][1][d])
        if body is not None:
            s.moving = 1
            s.body_m, s.body_k, s.body_c, s.body_added = body["m"], body["k"], body["c"], body["added_mass"]
            for k, key in enumerate(("d0", "v0", "a0")):
                vals = np.asarray(body.get(key, (0, 0)), dtype=np.float32)
                for i in range(2):
                    (s.d, s.v, s.a)[k][i] = float(vals[i])
        post = [(p[0], p[1]) for p in spec.get("post", ())]
        inlet = ("force_corrected_nebb", "left") if dim == 2 else ("nebb", "left")
        if post == [inlet, ("equilibrium", "right")]:
            kw_in, kw_out = spec["post"][0][2], spec["post"][1][2]
            if set(kw_in) != {"ux_wall"} or kw_in != kw_out:
                raise NotImplementedError("C port: inlet / outlet take ux_wall only")
            s.inlet_outlet, s.u0 = 1, float(kw_in["ux_wall"])
        elif post:
            raise NotImplementedError(f"C port: post list {post}")
        self.s = s
        self.f = np.ascontiguousarray(f0, dtype=dt).copy()
        ncell = int(np.prod(shape))
        self._tmp = np.empty_like(self.f)
        self._rho = np.empty(ncell, dtype=dt)
        self._u = np.empty(dim * ncell, dtype=dt)

    def run(self, n_steps, threads=0):
        rc = lib(self.dtype).ref_run(C.byref(self.s), self.f.ctypes.data_as(C.c_void_p), self._tmp.ctypes.data_as(C.c_void_p),
                           self._rho.ctypes.data_as(C.c_void_p), self._u.ctypes.data_as(C.c_void_p), int(n_steps),
                           int(threads))
        if rc != 0:
            raise RuntimeError("ref_run failed")
        return self.f

    def body_state(self):
        s = self.s
        return (np.array(s.d[:2], self.dtype), np.array(s.v[:2], self.dtype), np.array(s.a[:2], self.dtype),
                np.array(s.h[:2], self.dtype))
