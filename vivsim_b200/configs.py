"""Step descriptions for the benchmark configurations of BASELINE.json.

Each builder returns ``(spec, body)`` for ``vivsim_b200.Stepper``; ``body`` is None for a fixed
body.  Geometry follows the reference's fixtures: examples/benchmark.py:24-56 (C2),
examples/3d/flow_past_sphere.py:29-130 + examples/benchmark3d.py:30-105 (C3),
examples/2d/vortex_induced_vibration.py:33-74 (structural parameters)."""

import math

import numpy as np

from . import ib, ib3d


def get_omega(nu):
    return 1 / (3 * nu + 0.5)


def _window(markers, pad, headroom=0):
    lo = np.floor(markers.min(axis=0)).astype(int) - pad - headroom
    hi = np.floor(markers.max(axis=0)).astype(int) + pad + headroom + 1
    return tuple(int(x) for x in lo), tuple(int(x) for x in hi - lo)


def cavity(n=100, u0=0.5, nu=0.1):
    """C0/C1: README lid-driven cavity, BGK + NEE on four walls (README.md:89-122)."""
    spec = dict(dim=2, shape=(n, n), collision="bgk", omega=get_omega(nu), forcing=None,
                post=[("nee", "left", {}), ("nee", "right", {}), ("nee", "bottom", {}), ("nee", "top", {"ux_wall": u0})])
    return spec, None


def viv_cylinder_2d(nx=1024, ny=1024, n_marker=512, radius=50.0, u0=0.1, nu=0.01, n_iter=5,
                    collision="bgk", forcing="guo", moving=True, mass_ratio=10.0, reduced_velocity=5.0, pad=4,
                    center=None):
    """C2: D2Q9 IB-LBM cylinder, MDF + Guo forcing, inlet NEBB / outlet equilibrium, y periodic."""
    cx, cy = center if center is not None else (nx / 2, ny / 2)
    theta = np.linspace(0, 2 * np.pi, n_marker, endpoint=False).astype(np.float32)
    markers = np.stack([np.float32(cx) + np.float32(radius) * np.cos(theta),
                        np.float32(cy) + np.float32(radius) * np.sin(theta)], axis=1).astype(np.float32)
    headroom = 0   # the window follows the body (Stepper follow=1), so no extra room is needed
    origin, size = _window(markers, pad, headroom)
    spec = dict(dim=2, shape=(nx, ny), collision=collision, omega=get_omega(nu), forcing=forcing, u0=u0,
                ib=dict(markers=markers, ds=ib.get_ds(markers), kernel="peskin4", n_iter=n_iter, u_target=None,
                        window=(origin, size)),
                post=[("force_corrected_nebb", "left", {"ux_wall": u0}), ("equilibrium", "right", {"ux_wall": u0})])
    body = None
    if moving:
        d = 2 * radius
        area = math.pi * radius ** 2
        fn = u0 / (reduced_velocity * d)
        m = area * mass_ratio
        k = (2 * math.pi * fn) ** 2 * m * (1 + 1 / mass_ratio)
        body = dict(m=m, k=k, c=0.0, added_mass=area, n_dof=2, d0=(0.0, 0.0), v0=(0.0, 1e-2 * u0), a0=(0.0, 0.0))
    return spec, body


def sphere_3d(nx=256, ny=256, nz=256, diameter=48.0, u0=0.05, re=2000.0, n_iter=3, subdivisions=4,
              collision="kbc", forcing="edm", pad=4):
    """C3: D3Q19 flow past an immersed sphere (KBC, MDF 3 iterations, EDM, NEBB inlet / equilibrium outlet)."""
    verts, faces = icosphere(diameter / 2, (nx / 3.0, ny / 2.0, nz / 2.0), subdivisions)
    origin, size = _window(verts, pad)
    nu = u0 * diameter / re
    spec = dict(dim=3, shape=(nx, ny, nz), collision=collision, omega=get_omega(nu), forcing=forcing, u0=u0,
                ib=dict(markers=verts, ds=ib3d.get_ds(verts, faces), kernel="peskin4", n_iter=n_iter, u_target=None,
                        window=(origin, size)),
                post=[("nebb", "left", {"ux_wall": u0}), ("equilibrium", "right", {"ux_wall": u0})])
    return spec, None


def viv_cylinder_2d_large(n=16384, u0=0.05, re=1e4, n_iter=5, center_x=None):
    """C4: D2Q9 KBC VIV cylinder Re = 1e4 on n x n, D = n/20, 4D markers, EDM.  center_x defaults to 3n/16, the
    middle of the second of eight slabs, so the same geometry runs on 1, 2, 4 and 8 GPUs."""
    d = n / 20
    cx = 3 * n / 16 if center_x is None else center_x
    spec, body = viv_cylinder_2d(nx=n, ny=n, n_marker=int(4 * d), radius=d / 2, u0=u0, nu=u0 * d / re, n_iter=n_iter,
                                 collision="kbc", forcing="edm", moving=True, center=(cx, n / 2))
    return spec, body


def cylinder_surface_markers(center, radius, height, spacing=0.5):
    """Markers and area weights of a finite z-aligned cylinder: lateral surface on a (theta, z) grid and two caps of
    concentric rings, all at <= `spacing` lattice units (the resolution rule of examples/3d/oscillating_cylinder.py:
    146-149).  The weight of a marker is the surface area it represents."""
    cx, cy, cz = center
    n_t = int(math.ceil(2 * math.pi * radius / spacing))
    n_z = int(math.ceil(height / spacing))
    theta = np.linspace(0.0, 2 * math.pi, n_t, endpoint=False)
    z = np.linspace(cz - height / 2, cz + height / 2, n_z + 1)
    tt, zz = np.meshgrid(theta, z, indexing="ij")
    lat = np.stack([cx + radius * np.cos(tt), cy + radius * np.sin(tt), zz], axis=-1).reshape(-1, 3)
    wz = np.full(n_z + 1, height / n_z)
    wz[0] = wz[-1] = 0.5 * height / n_z
    w_lat = ((2 * math.pi * radius / n_t) * np.broadcast_to(wz, (n_t, n_z + 1))).reshape(-1)
    pts, wts = [lat], [w_lat]
    n_r = max(1, int(math.ceil(radius / spacing)))
    edges = np.linspace(0.0, radius, n_r + 1)
    for z_cap in (z[0], z[-1]):
        for k in range(n_r):                     # ring k represents the annulus edges[k] .. edges[k + 1]
            r_mid = 0.5 * (edges[k] + edges[k + 1])
            n_k = max(6, int(math.ceil(2 * math.pi * r_mid / spacing)))
            t = np.linspace(0.0, 2 * math.pi, n_k, endpoint=False)
            pts.append(np.stack([cx + r_mid * np.cos(t), cy + r_mid * np.sin(t), np.full(n_k, z_cap)], axis=-1))
            wts.append(np.full(n_k, math.pi * (edges[k + 1] ** 2 - edges[k] ** 2) / n_k))
    return np.concatenate(pts).astype(np.float32), np.concatenate(wts).astype(np.float32)


def oscillating_cylinder_3d(nx=1024, ny=512, nz=512, diameter=None, u0=0.05, re=1000.0, n_iter=3, collision="mrt",
                            forcing="guo", mass_ratio=2.0, reduced_velocity=5.0, center_x=None, pad=4, moving=True):
    """C5: D3Q19 MRT flow past an elastically mounted finite cylinder along z (examples/3d/oscillating_cylinder.py:
    229-282 recipe): MDF(3), Guo-MRT forcing, NEBB inlet / equilibrium outlet, y and z periodic, 2-DOF Newmark body,
    IB window following the body with clip(floor()).  center_x defaults to 5 nx / 16 (middle of the third of eight
    slabs)."""
    d = nx / 10 if diameter is None else diameter
    h = nz - 24
    cx = 5 * nx / 16 if center_x is None else center_x
    center = (cx, ny / 2, nz / 2)
    markers, ds = cylinder_surface_markers(center, d / 2, h)
    origin, size = _window(markers, pad)
    nu = u0 * d / re
    spec = dict(dim=3, shape=(nx, ny, nz), collision=collision, omega=get_omega(nu), forcing=forcing, u0=u0,
                ib=dict(markers=markers, ds=ds, kernel="peskin4", n_iter=n_iter, u_target=None, window=(origin, size)),
                post=[("nebb", "left", {"ux_wall": u0}), ("equilibrium", "right", {"ux_wall": u0})])
    body = None
    if moving:
        vol = math.pi * (d / 2) ** 2 * h
        fn = u0 / (reduced_velocity * d)
        m = vol * mass_ratio
        k = (2 * math.pi * fn) ** 2 * m * (1 + 1 / mass_ratio)
        body = dict(m=m, k=k, c=0.0, added_mass=vol, n_dof=2, d0=(0.0, 0.0), v0=(0.0, 1e-2 * u0), a0=(0.0, 0.0))
    return spec, body


def icosphere(radius, center, subdivisions):
    """Unit icosahedron subdivided `subdivisions` times: 10 * 4^n + 2 vertices (fixture generator)."""
    phi = (1 + 5 ** 0.5) / 2
    base = [(-1, phi, 0), (1, phi, 0), (-1, -phi, 0), (1, -phi, 0), (0, -1, phi), (0, 1, phi),
            (0, -1, -phi), (0, 1, -phi), (phi, 0, -1), (phi, 0, 1), (-phi, 0, -1), (-phi, 0, 1)]
    verts = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in base]
    tris = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2),
            (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11),
            (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    for _ in range(subdivisions):
        cache, nxt = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]

        for a, b, c in tris:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nxt += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        tris = nxt
    pts = np.array(verts) * radius + np.asarray(center, dtype=np.float64)
    return pts.astype(np.float32), np.array(tris, dtype=np.int32)


def uniform_state(spec, noise=0.0, seed=0):
    """f = feq(rho = 1, u = (u0, 0[, 0]) + noise N(0,1)) as a CUDA tensor, computed by the library itself."""
    import torch
    from . import lbm, lbm3d
    shape, dim = tuple(spec["shape"]), spec["dim"]
    u = torch.zeros((dim,) + shape, device="cuda")
    u[0] = float(spec.get("u0", 0.0))
    if noise:
        gen = torch.Generator(device="cuda").manual_seed(seed)
        u += noise * torch.randn(u.shape, device="cuda", generator=gen)
    mod = lbm if dim == 2 else lbm3d
    return mod.get_equilibrium(torch.ones(shape, device="cuda"), u)
