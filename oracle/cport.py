"""ctypes wrapper of oracle/c/iblbm_ref.c (C + OpenMP restatement of the composed step).
TEST INFRASTRUCTURE ONLY: a fast checker for long / large runs and bench.py's CPU baseline.

Covers the recipes of BASELINE configs 1-3: collision bgk|kbc|reg, forcing none|edm|guo, uniform g,
one immersed body (Peskin 4-point, MDF, optional 2-DOF Newmark), inlet NEBB / outlet equilibrium on
the x faces or fully periodic.  Anything else raises."""

import ctypes as C
import os

import numpy as np

from .c import build as _build

_COLL = {"bgk": 0, "kbc": 2, "reg": 3}
_FORCE = {None: 0, "edm": 1, "guo": 2}


class RefSpec(C.Structure):
    _fields_ = [("dim", C.c_int), ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("collision", C.c_int),
                ("forcing", C.c_int), ("omega", C.c_double), ("g0", C.c_float * 3), ("n_markers", C.c_int),
                ("n_iter", C.c_int), ("markers0", C.c_void_p), ("ds", C.c_void_p), ("worg0", C.c_float * 3),
                ("wsz", C.c_int * 3), ("moving", C.c_int), ("body_m", C.c_double), ("body_k", C.c_double),
                ("body_c", C.c_double), ("body_added", C.c_double), ("d", C.c_float * 3), ("v", C.c_float * 3),
                ("a", C.c_float * 3), ("h", C.c_float * 3), ("inlet_outlet", C.c_int), ("u0", C.c_float),
                ("marker_force", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.ref_num_threads.restype = C.c_int
    return _lib


def num_threads():
    return int(lib().ref_num_threads())


class CRunner:
    """Holds the buffers of one simulation so repeated ``run`` calls can be timed without allocation."""

    def __init__(self, spec, f0, body=None):
        dim = spec["dim"]
        shape = tuple(spec["shape"])
        s = RefSpec()
        s.dim, s.nx, s.ny, s.nz = dim, shape[0], shape[1], (shape[2] if dim == 3 else 1)
        if spec["collision"] not in _COLL:
            raise NotImplementedError(f"C port: collision {spec['collision']!r}")
        s.collision, s.forcing, s.omega = _COLL[spec["collision"]], _FORCE[spec.get("forcing")], float(spec["omega"])
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
            self._markers = np.ascontiguousarray(ib["markers"], dtype=np.float32)
            m = self._markers.shape[0]
            self._ds = np.ascontiguousarray(np.broadcast_to(np.asarray(ib["ds"], dtype=np.float32), (m,)))
            self.marker_force = np.zeros((m, dim), dtype=np.float32)
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
        self.f = np.ascontiguousarray(f0, dtype=np.float32).copy()
        ncell = int(np.prod(shape))
        self._tmp = np.empty_like(self.f)
        self._rho = np.empty(ncell, dtype=np.float32)
        self._u = np.empty(dim * ncell, dtype=np.float32)

    def run(self, n_steps, threads=0):
        rc = lib().ref_run(C.byref(self.s), self.f.ctypes.data_as(C.c_void_p), self._tmp.ctypes.data_as(C.c_void_p),
                           self._rho.ctypes.data_as(C.c_void_p), self._u.ctypes.data_as(C.c_void_p), int(n_steps),
                           int(threads))
        if rc != 0:
            raise RuntimeError("ref_run failed")
        return self.f

    def body_state(self):
        s = self.s
        return (np.array(s.d[:2], np.float32), np.array(s.v[:2], np.float32), np.array(s.a[:2], np.float32),
                np.array(s.h[:2], np.float32))
