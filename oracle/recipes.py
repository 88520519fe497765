"""Composed IB-LBM time steps (oracle; NumPy fp32; test infrastructure only).

A step is described by a plain dict (``spec``) that the product's fused
``vivsim_b200.Stepper`` accepts as well, so parity tests feed one description to
both sides.  The sequence of calls follows the reference's example drivers:

* README.md:104-122 / examples/2d/lid_driven_cavity.py:55-64   (collide, stream, BCs in order)
* examples/2d/poiseuille_channel.py:80-148                     (EDM: uncorrected u; Guo: u += g/2rho first)
* examples/2d/flow_pass_cylinder.py:97-125, examples/3d/flow_past_sphere.py:139-171 (fixed body IB, EDM)
* examples/2d/vortex_induced_vibration.py:96-148               (moving body, Newmark on the total marker force)
* examples/2d/flow_through_text.py:70-79                        (obstacle mask after the BCs)

spec keys
---------
dim        2 | 3
collision  "bgk" | "mrt" | "kbc" | "reg";  omega  float
forcing    None | "edm" | "guo"
g          None | sequence of dim floats (uniform body force) | array (dim, *shape)
ib         None | dict(markers (M,dim) absolute coords, ds scalar|(M,), kernel "peskin4"|"peskin3"|"cosine4"|"hat2",
                       n_iter int, u_target None|(M,dim), window (origin tuple, size tuple))
post       ordered list of ("<bc name without 'boundary_'>", loc, kwargs) and ("mask", mask_array)
"""

import numpy as np

from . import ib as ib2, ib3d, lbm, lbm3d, dyn
from .core import F32, f32

KERNELS = {"peskin4": ib2.kernel_peskin_4pt, "peskin3": ib2.kernel_peskin_3pt,
           "cosine4": ib2.kernel_cosine_4pt, "hat2": ib2.kernel_hat_2pt}


def _mod(spec):
    return lbm if spec["dim"] == 2 else lbm3d


def _window_slices(window):
    origin, size = window
    return tuple(slice(int(o), int(o) + int(n)) for o, n in zip(origin, size))


def ib_force(spec, u, markers=None, u_target=None):
    """MDF on the IB window: returns (g on the full grid, marker reaction force (M,dim)).

    Follows the window pattern of flow_pass_cylinder.py:104-121 (dynamic_slice of u,
    stencil in window coordinates, dynamic_update_slice back)."""
    ibs = spec["ib"]
    dim = spec["dim"]
    markers = f32(ibs["markers"] if markers is None else markers)
    win = _window_slices(ibs["window"])
    origin = np.asarray(ibs["window"][0], dtype=F32)
    size = tuple(int(n) for n in ibs["window"][1])
    local = (markers - origin).astype(F32)
    kern = KERNELS[ibs.get("kernel", "peskin4")]
    if dim == 2:
        w, idx = ib2.get_ib_stencil(local[:, 0], local[:, 1], size[1], kernel=kern)
    else:
        w, idx = ib3d.get_ib_stencil(local, size, kernel=kern)
    tgt = u_target if u_target is not None else ibs.get("u_target")
    tgt = np.zeros_like(markers) if tgt is None else f32(tgt)
    u_win = u[(slice(None),) + win]
    g_win, h = ib2.multi_direct_forcing(u_win, w, idx, tgt, ibs["ds"], n_iter=ibs.get("n_iter", 5))
    g = np.zeros_like(u)
    g[(slice(None),) + win] = g_win
    return g, h


def _body_force(spec, shape):
    g = spec.get("g")
    if g is None:
        return None
    g = f32(g)
    if g.ndim == 1:
        return np.broadcast_to(g.reshape((-1,) + (1,) * len(shape)), (spec["dim"],) + tuple(shape)).astype(F32)
    return g


def collide(spec, f, markers=None, u_target=None):
    """moments -> (IB force) -> collision -> forcing.  Returns (f_post, rho, u, marker_force|None)."""
    m = _mod(spec)
    f = f32(f)
    rho, u = m.get_macroscopic(f)
    g = _body_force(spec, rho.shape)
    h = None
    if spec.get("ib") is not None:
        g_ib, h = ib_force(spec, u, markers, u_target)
        g = g_ib if g is None else (g + g_ib).astype(F32)
    forcing = spec.get("forcing")
    if forcing == "guo" and g is not None:
        u = (u + m.get_velocity_correction(g, rho)).astype(F32)
    feq = m.get_equilibrium(rho, u)
    kind, omega = spec["collision"], spec["omega"]
    if kind == "bgk":
        f = m.collision_bgk(f, feq, omega)
    elif kind == "kbc":
        f = m.collision_kbc(f, feq, omega)
    elif kind == "reg":
        f = m.collision_reg(f, feq, omega)
    elif kind == "mrt":
        f = m.collision_mrt(f, feq, spec.get("mrt_op", None) if spec.get("mrt_op") is not None
                            else m.get_mrt_collision_operator(omega))
    else:
        raise ValueError(kind)
    if g is not None and forcing is not None:
        if forcing == "edm":
            f = m.forcing_edm(f, g, u)
        elif forcing == "guo" and kind == "mrt":
            f = m.forcing_guo_mrt(f, g, u, spec.get("mrt_fop", None) if spec.get("mrt_fop") is not None
                                  else m.get_mrt_forcing_operator(omega))
        elif forcing == "guo":
            f = m.forcing_guo_bgk(f, g, u, omega)
        else:
            raise ValueError(forcing)
    return f, rho, u, h


def stream_and_post(spec, f_post):
    """streaming followed by the ordered post-stream operations."""
    m = _mod(spec)
    f = m.streaming(f_post)
    for op in spec.get("post", ()):
        if op[0] == "mask":
            f = m.obstacle_bounce_back(f, op[1])
        elif op[0] in ("bounce_back", "specular_reflection"):
            f = getattr(m, "boundary_" + op[0])(f_post, f, op[1], **(op[2] if len(op) > 2 else {}))
        else:
            f = getattr(m, "boundary_" + op[0])(f, op[1], **(op[2] if len(op) > 2 else {}))
    return f


def step(spec, f, markers=None, u_target=None):
    """One reference time step on the reference's state convention (post-BC f)."""
    f_post, _, _, h = collide(spec, f, markers, u_target)
    return stream_and_post(spec, f_post), h


def run(spec, f, n_steps):
    h = None
    for _ in range(n_steps):
        f, h = step(spec, f)
    return f, h


def viv_step(spec, body, f, d, v, a, follow=1):
    """Moving rigid body coupled through Newmark-beta (2-DOF translation).

    follow = 1: window origin trunc(origin0 + d) (vortex_induced_vibration.py:104-105); follow = 2:
    clip(floor(origin0 + d), 0, N - size) (examples/3d/oscillating_cylinder.py:241-243).

    ``body`` = dict(m, k, c, added_mass) ; markers in ``spec['ib']['markers']`` are the
    initial coordinates, shifted by ``d`` every step; the IB window origin follows
    trunc(origin0 + d) as in vortex_induced_vibration.py:104-105,112-137."""
    dim = spec["dim"]
    ibs = dict(spec["ib"])
    d = f32(d); v = f32(v); a = f32(a)
    origin0, size = ibs["window"]
    origin = list(origin0)
    for k in range(len(d)):
        shifted = F32(origin0[k]) + d[k]
        if follow == 2:
            origin[k] = int(min(max(int(np.floor(shifted)), 0), spec["shape"][k] - size[k]))
        else:
            origin[k] = int(np.trunc(shifted))
    ibs["window"] = (tuple(origin), size)
    shift = np.zeros(dim, dtype=F32); shift[:len(d)] = d
    markers = (f32(spec["ib"]["markers"]) + shift).astype(F32)
    tgt = np.zeros_like(markers); tgt[:, :len(v)] = v
    sp = dict(spec); sp["ib"] = ibs
    f_post, _, _, hm = collide(sp, f, markers, tgt)
    h = dyn.get_force_to_obj(hm)[:len(d)] + a * F32(body["added_mass"])
    a2, v2, d2 = dyn.newmark_2dof(a, v, d, h, body["m"], body["k"], body["c"])
    return stream_and_post(sp, f_post), d2, v2, a2, h


def viv_step_3dof(spec, body, f, d, v, a, follow=1):
    """Moving rigid body with three degrees of freedom in 2-D (x, y, rotation about body['center']).

    Marker coordinates and target velocities from dyn.py:84-120, total force and torque from dyn.py:123-154, matrix-form
    Newmark (dyn.py:36-42); the rest as viv_step.  ``body`` = dict(m, k, c (3 x 3), added_mass (3,), center)."""
    ibs = dict(spec["ib"])
    d = f32(d); v = f32(v); a = f32(a)
    origin0, size = ibs["window"]
    origin = list(origin0)
    for k in range(2):
        shifted = F32(origin0[k]) + d[k]
        if follow == 2:
            origin[k] = int(min(max(int(np.floor(shifted)), 0), spec["shape"][k] - size[k]))
        else:
            origin[k] = int(np.trunc(shifted))
    ibs["window"] = (tuple(origin), size)
    xc, yc = body["center"]
    m0 = f32(spec["ib"]["markers"])
    mx, my = dyn.get_markers_coords_3dof(m0[:, 0], m0[:, 1], xc, yc, d)
    markers = np.stack([mx, my], axis=1).astype(F32)
    tgt = dyn.get_markers_velocity_3dof(mx, my, xc, yc, d, v)
    sp = dict(spec); sp["ib"] = ibs
    f_post, _, _, hm = collide(sp, f, markers, tgt)
    h = np.concatenate([dyn.get_force_to_obj(hm), [dyn.get_torque_to_obj(mx, my, xc, yc, d, hm)]]).astype(F32)
    h = (h + a * f32(body["added_mass"])).astype(F32)
    a2, v2, d2 = dyn.newmark_3dof(a, v, d, h, f32(body["m"]), f32(body["k"]), f32(body["c"]))
    return stream_and_post(sp, f_post), d2, v2, a2, h


# ------------------------------------------------------------------ named configurations
def cavity_spec(n=100, u0=0.5, nu=0.1):
    """BASELINE config 0 (README.md:89-122): BGK, NEE on four walls, lid last."""
    return dict(dim=2, shape=(n, n), collision="bgk", omega=lbm.get_omega(nu), forcing=None,
                post=[("nee", "left", {}), ("nee", "right", {}), ("nee", "bottom", {}),
                      ("nee", "top", {"ux_wall": u0})])


def cavity_init(spec):
    rho = np.ones(spec["shape"], dtype=F32)
    return lbm.get_equilibrium(rho, np.zeros((2,) + tuple(spec["shape"]), dtype=F32))


def cylinder2d_spec(nx=1024, ny=1024, n_marker=512, radius=50.0, u0=0.1, nu=0.01, n_iter=5,
                    collision="bgk", forcing="guo", pad=4):
    """BASELINE config 1 (examples/benchmark.py:24-56 fixtures; SURVEY 8d C2 recipe)."""
    theta = np.linspace(0, 2 * np.pi, n_marker, endpoint=False).astype(F32)
    mx = (F32(nx / 2) + F32(radius) * np.cos(theta)).astype(F32)
    my = (F32(ny / 2) + F32(radius) * np.sin(theta)).astype(F32)
    markers = np.stack([mx, my], axis=1)
    lo = np.floor(markers.min(axis=0)).astype(int) - pad
    hi = np.floor(markers.max(axis=0)).astype(int) + pad + 1
    return dict(dim=2, shape=(nx, ny), collision=collision, omega=lbm.get_omega(nu), forcing=forcing,
                u0=u0,
                ib=dict(markers=markers, ds=ib2.get_ds(markers), kernel="peskin4", n_iter=n_iter,
                        u_target=None, window=(tuple(int(x) for x in lo), tuple(int(x) for x in hi - lo))),
                post=[("force_corrected_nebb", "left", {"ux_wall": u0}),
                      ("equilibrium", "right", {"ux_wall": u0})])


def sphere3d_spec(nx=256, ny=256, nz=256, diameter=48.0, u0=0.05, re=2000.0, n_iter=3, subdivisions=4,
                  collision="kbc", forcing="edm", pad=4):
    """BASELINE config 2 (examples/3d/flow_past_sphere.py:139-171 recipe at benchmark3d.py shape)."""
    center = (nx / 3.0, ny / 2.0, nz / 2.0)
    verts, faces = ib3d.icosphere(diameter / 2, center, subdivisions)
    lo = np.floor(verts.min(axis=0)).astype(int) - pad
    hi = np.floor(verts.max(axis=0)).astype(int) + pad + 1
    nu = u0 * diameter / re
    return dict(dim=3, shape=(nx, ny, nz), collision=collision, omega=lbm.get_omega(nu), forcing=forcing,
                u0=u0,
                ib=dict(markers=verts, ds=ib3d.get_ds(verts, faces), kernel="peskin4", n_iter=n_iter,
                        u_target=None, window=(tuple(int(x) for x in lo), tuple(int(x) for x in hi - lo))),
                post=[("nebb", "left", {"ux_wall": u0}), ("equilibrium", "right", {"ux_wall": u0})])


def uniform_init(spec, noise=0.0, seed=0):
    """f = feq(1, (u0,0[,0]) + noise*N(0,1)) (perturbed equilibrium, seed 0)."""
    shape = tuple(spec["shape"])
    dim = spec["dim"]
    u = np.zeros((dim,) + shape, dtype=F32)
    u[0] = F32(spec.get("u0", 0.0))
    if noise:
        u += F32(noise) * np.random.default_rng(seed).standard_normal(u.shape).astype(F32)
    return _mod(spec).get_equilibrium(np.ones(shape, dtype=F32), u)
