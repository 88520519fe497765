"""Lattice-generic implementation of the reference's pure-function API on CUDA tensors.

``vivsim_b200.lbm`` (D2Q9) and ``vivsim_b200.lbm3d`` (D3Q19) bind these with the reference's
names and signatures (vivsim/lbm/__init__.py:1-42, vivsim/lbm3d/__init__.py:1-44).  Arrays are
torch CUDA tensors (fp32); every function returns new tensors like the reference's jnp functions.
"""

import ctypes as C

import numpy as np
import torch

from . import _lib as L

Q = {2: 9, 3: 19}

# D2Q9 / D3Q19 moment bases (reference lbm/collision/mrt.py:10-22, lbm3d/collision/mrt.py:7-30),
# generated from polynomials of the lattice velocities.
_C2 = np.array([[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1], [1, 1], [-1, 1], [-1, -1], [1, -1]], dtype=np.float64)
_C3 = np.array([[0, 0, 0], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1],
                [1, 1, 0], [-1, 1, 0], [1, -1, 0], [-1, -1, 0], [1, 0, 1], [-1, 0, 1], [1, 0, -1], [-1, 0, -1],
                [0, 1, 1], [0, -1, 1], [0, 1, -1], [0, -1, -1]], dtype=np.float64)


def _basis(dim):
    if dim == 2:
        x, y = _C2.T
        e = x * x + y * y
        return np.array([np.ones(9), 3 * e - 4, 4 - 10.5 * e + 4.5 * e * e, x, x * (3 * e - 5), y, y * (3 * e - 5),
                         x * x - y * y, x * y])
    x, y, z = _C3.T
    return np.array([np.ones(19), x, y, z, x * x + y * y + z * z, 2 * x * x - y * y - z * z, y * y - z * z,
                     x * y, x * z, y * z, x * x * y, x * x * z, x * y * y, y * y * z, x * z * z, y * z * z,
                     x * x * y * y, x * x * z * z, y * y * z * z])


def mrt_rates(dim, omega):
    """diag(S): lbm/collision/mrt.py:44, lbm3d/collision/mrt.py:50-72."""
    if dim == 2:
        return [0, 1.4, 1.4, 0, 1.2, 0, 1.2, omega, omega]
    return [0, 0, 0, 0, 1.1] + [omega] * 5 + [1.2] * 6 + [1.4] * 3


def mrt_operator(dim, omega, forcing=False):
    """M^-1 S M or M^-1 (I - S/2) M as an fp32 host matrix (float64 products, one rounding)."""
    M = _basis(dim)
    s = np.asarray(mrt_rates(dim, float(omega)), dtype=np.float64)
    core = np.diag(1.0 - 0.5 * s) if forcing else np.diag(s)
    return (np.linalg.inv(M) @ core @ M).astype(np.float32)


def get_omega(nu):
    """omega = 1 / (3 nu + 0.5)   (lbm/basic.py:159-172)."""
    return 1 / (3 * nu + 0.5)


def get_velocity_correction(g, rho=1):
    """g / (2 rho)   (lbm/basic.py:175-194).  Tiny elementwise helper kept in torch."""
    return g * 0.5 / rho


def _field(dim, f, lead, name):
    """Check a (lead, *spatial) CUDA field; spatial may be any shape (edge slices allowed)."""
    f = L.dev(f, name=name)
    if f.ndim < 1 or f.shape[0] != lead:
        raise ValueError(f"{name}: leading axis must be {lead}, got shape {tuple(f.shape)}")
    return f


def streaming(dim, f):
    f = _field(dim, f, Q[dim], "f")
    if f.ndim != dim + 1:
        raise ValueError(f"streaming: f must have shape ({Q[dim]}, {'NX, NY' if dim == 2 else 'NX, NY, NZ'})")
    out = torch.empty_like(f)
    grid = L.grid_of(f.shape[1:])
    L.check(L.lib().vsb_streaming(C.byref(grid), L.ptr(f), L.ptr(out), L.stream()))
    return out


def get_macroscopic(dim, f):
    f = _field(dim, f, Q[dim], "f")
    spatial = tuple(f.shape[1:])
    n = int(np.prod(spatial)) if spatial else 1
    rho = torch.empty(spatial, device=f.device, dtype=torch.float32)
    u = torch.empty((dim,) + spatial, device=f.device, dtype=torch.float32)
    L.check(L.lib().vsb_macroscopic(dim, C.c_int64(n), L.ptr(f), L.ptr(rho), L.ptr(u), L.stream()))
    return rho, u


def get_equilibrium(dim, rho, u):
    u = _field(dim, u, dim, "u")
    rho = L.dev(rho if isinstance(rho, torch.Tensor) else torch.as_tensor(rho, device=u.device), name="rho")
    spatial = tuple(u.shape[1:])
    if tuple(rho.shape) != spatial:
        rho = rho.expand(spatial).contiguous()
    n = int(np.prod(spatial)) if spatial else 1
    feq = torch.empty((Q[dim],) + spatial, device=u.device, dtype=torch.float32)
    L.check(L.lib().vsb_equilibrium(dim, C.c_int64(n), L.ptr(rho), L.ptr(u), L.ptr(feq), L.stream()))
    return feq


def _collision(dim, kind, f, feq, omega=1.0, op=None):
    f = _field(dim, f, Q[dim], "f")
    feq = _field(dim, feq, Q[dim], "feq")
    if f.shape != feq.shape:
        raise ValueError(f"f {tuple(f.shape)} and feq {tuple(feq.shape)} differ")
    out = torch.empty_like(f)
    op_h = L.host_matrix(op, Q[dim]) if op is not None else None
    L.check(L.lib().vsb_collision(dim, C.c_int64(f[0].numel()), L.COLL[kind], C.c_double(float(omega)),
                                  op_h.ctypes.data_as(C.c_void_p) if op_h is not None else None,
                                  L.ptr(f), L.ptr(feq), L.ptr(out), L.stream()))
    return out


def collision_bgk(dim, f, feq, omega): return _collision(dim, "bgk", f, feq, omega)
def collision_kbc(dim, f, feq, omega): return _collision(dim, "kbc", f, feq, omega)
def collision_reg(dim, f, feq, omega): return _collision(dim, "reg", f, feq, omega)
def collision_mrt(dim, f, feq, mrt_collision_matrix): return _collision(dim, "mrt", f, feq, 1.0, mrt_collision_matrix)


def get_guo_forcing_term(dim, g, u):
    g = _field(dim, g, dim, "g")
    u = _field(dim, u, dim, "u")
    if g.shape != u.shape:
        raise ValueError(f"g {tuple(g.shape)} and u {tuple(u.shape)} differ")
    out = torch.empty((Q[dim],) + tuple(u.shape[1:]), device=u.device, dtype=torch.float32)
    L.check(L.lib().vsb_guo_forcing_term(dim, C.c_int64(u[0].numel()), L.ptr(g), L.ptr(u), L.ptr(out), L.stream()))
    return out


def _forcing(dim, kind, f, g, u, omega=1.0, fop=None):
    f = _field(dim, f, Q[dim], "f")
    g = _field(dim, g, dim, "g")
    u = _field(dim, u, dim, "u")
    if g.shape != u.shape or f.shape[1:] != u.shape[1:]:
        raise ValueError("f, g, u spatial shapes differ")
    out = torch.empty_like(f)
    fop_h = L.host_matrix(fop, Q[dim]) if fop is not None else None
    L.check(L.lib().vsb_forcing(dim, C.c_int64(f[0].numel()), L.FORCE[kind], C.c_double(float(omega)),
                                fop_h.ctypes.data_as(C.c_void_p) if fop_h is not None else None,
                                L.ptr(f), L.ptr(g), L.ptr(u), L.ptr(out), L.stream()))
    return out


def forcing_edm(dim, f, g, u): return _forcing(dim, "edm", f, g, u)
def forcing_guo_bgk(dim, f, g, u, omega): return _forcing(dim, "guo", f, g, u, omega)
def forcing_guo_mrt(dim, f, g, u, mrt_forcing_operator): return _forcing(dim, "guo", f, g, u, 1.0, mrt_forcing_operator)


# ---------------------------------------------------------------------------- post-streaming operations
_COMP = ("ux_wall", "uy_wall", "uz_wall")
_GCOMP = ("gx_wall", "gy_wall", "gz_wall")


def face_shape(dim, shape, loc):
    if loc not in L.LOC or L.LOC[loc] >= 2 * dim:
        raise KeyError(loc)  # the reference's dict lookup raises KeyError for a bad loc
    axis = L.LOC[loc] // 2
    return tuple(n for a, n in enumerate(shape) if a != axis)


def make_post_op(dim, shape, kind, loc=None, wrap="", keep=None, mask=None, **kw):
    """Build a VsbPostOp from reference-style keyword arguments (rho_wall, ux_wall, ..., gx_wall, ...)."""
    keep = [] if keep is None else keep
    op = L.VsbPostOp()
    op.kind = L.BC[kind]
    if kind == "mask":
        m = mask
        if not isinstance(m, torch.Tensor) or not m.is_cuda:
            raise L.VsbError("mask: expected a CUDA bool / uint8 tensor")
        if tuple(m.shape) != tuple(shape):
            raise ValueError(f"mask shape {tuple(m.shape)} != grid shape {tuple(shape)}")
        m = m.to(torch.uint8).contiguous()
        keep.append(m)
        op.mask = m.data_ptr()
        return op, keep
    fs = face_shape(dim, shape, loc)
    op.wrap = L.WRAP[wrap]
    op.loc = L.LOC[loc]
    allowed = {"rho_wall"} | set(_COMP[:dim]) | set(_GCOMP[:dim])
    bad = set(kw) - allowed
    if bad:
        raise TypeError(f"unexpected keyword argument(s) {sorted(bad)}")
    op.rho = L.wall_value(kw.get("rho_wall", 1), fs, keep, "rho_wall")
    for d in range(dim):
        op.u[d] = L.wall_value(kw.get(_COMP[d], 0), fs, keep, _COMP[d])
        op.g[d] = L.wall_value(kw.get(_GCOMP[d], 0), fs, keep, _GCOMP[d])
    return op, keep


def post_op(dim, f, kind, loc=None, wrap="", f_before_stream=None, mask=None, inplace=False, **kw):
    f = _field(dim, f, Q[dim], "f")
    if f.ndim != dim + 1:
        raise ValueError(f"f must have {dim} spatial axes")
    out = f if inplace else f.clone()
    pre = None
    if f_before_stream is not None:
        pre = _field(dim, f_before_stream, Q[dim], "f_before_stream")
        if pre.shape != f.shape:
            raise ValueError("f_before_stream and f differ in shape")
    op, keep = make_post_op(dim, f.shape[1:], kind, loc, wrap, mask=mask, **kw)
    grid = L.grid_of(f.shape[1:])
    L.check(L.lib().vsb_post_op(C.byref(grid), C.byref(op), L.ptr(pre), L.ptr(out), L.stream()))
    return out


def boundary_characteristic(dim, rho, u, loc="right"):
    if loc not in L.LOC or L.LOC[loc] >= 2 * dim:
        raise ValueError("loc must name a face of the lattice")
    u = _field(dim, u, dim, "u")
    rho = L.dev(rho, name="rho")
    fs = face_shape(dim, rho.shape, loc)
    rho_out = torch.empty(fs, device=u.device, dtype=torch.float32)
    u_out = torch.empty((dim,) + fs, device=u.device, dtype=torch.float32)
    grid = L.grid_of(rho.shape)
    L.check(L.lib().vsb_boundary_characteristic(C.byref(grid), L.LOC[loc], L.ptr(rho), L.ptr(u), L.ptr(rho_out),
                                                L.ptr(u_out), L.stream()))
    return rho_out, u_out


def bind(dim, namespace):
    """Populate a module namespace with the reference's function names for one lattice."""
    comps = _COMP[:dim]
    gcomps = _GCOMP[:dim]

    def fix(fn):
        def wrapped(*a, **k):
            return fn(dim, *a, **k)
        wrapped.__name__ = fn.__name__
        wrapped.__doc__ = fn.__doc__
        return wrapped

    for fn in (streaming, get_macroscopic, get_equilibrium, collision_bgk, collision_kbc, collision_reg, collision_mrt,
               get_guo_forcing_term, forcing_edm, forcing_guo_bgk, forcing_guo_mrt, boundary_characteristic):
        namespace[fn.__name__] = fix(fn)
    namespace["get_omega"] = get_omega
    namespace["get_velocity_correction"] = get_velocity_correction
    namespace["get_mrt_collision_operator"] = lambda omega: mrt_operator(dim, omega)
    namespace["get_mrt_forcing_operator"] = lambda omega: mrt_operator(dim, omega, forcing=True)

    def make_core(kind):
        def core(f, loc, rho_wall=1, **kw):
            return post_op(dim, f, kind, loc, rho_wall=rho_wall, **kw)
        core.__name__ = f"boundary_{kind}"
        return core

    def make_velocity(kind):
        def fn(f, loc, **kw):
            if "rho_wall" in kw:
                raise TypeError("velocity boundaries derive rho_wall from the populations")
            return post_op(dim, f, kind, loc, wrap="velocity", **kw)
        fn.__name__ = f"boundary_velocity_{kind}"
        return fn

    def make_pressure(kind):
        def fn(f, loc, rho_wall=1):
            return post_op(dim, f, kind, loc, wrap="pressure", rho_wall=rho_wall)
        fn.__name__ = f"boundary_pressure_{kind}"
        return fn

    def make_force_corrected(kind):
        def fn(f, loc, rho_wall=1, **kw):
            return post_op(dim, f, kind, loc, wrap="force_corrected", rho_wall=rho_wall, **kw)
        fn.__name__ = f"boundary_force_corrected_{kind}"
        return fn

    for kind in ("nee", "nebb", "equilibrium"):
        namespace[f"boundary_{kind}"] = make_core(kind)
        namespace[f"boundary_velocity_{kind}"] = make_velocity(kind)
        namespace[f"boundary_pressure_{kind}"] = make_pressure(kind)
        namespace[f"boundary_force_corrected_{kind}"] = make_force_corrected(kind)

    def boundary_bounce_back(f_before_stream, f, loc, **kw):
        return post_op(dim, f, "bounce_back", loc, f_before_stream=f_before_stream, **kw)

    def boundary_specular_reflection(f_before_stream, f, loc, **kw):
        return post_op(dim, f, "specular_reflection", loc, f_before_stream=f_before_stream, **kw)

    def obstacle_bounce_back(f, mask):
        return post_op(dim, f, "mask", mask=mask)

    namespace["boundary_bounce_back"] = boundary_bounce_back
    namespace["boundary_specular_reflection"] = boundary_specular_reflection
    namespace["obstacle_bounce_back"] = obstacle_bounce_back
    namespace["__all__"] = [k for k in namespace if not k.startswith("_")]
    del comps, gcomps
