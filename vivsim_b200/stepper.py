"""Fused IB-LBM time stepper: the hot path.

The reference composes a step from separately-jitted functions (e.g.
examples/2d/vortex_induced_vibration.py:96-148):

    moments -> equilibrium -> collision -> [IB: window, stencil, multi-direct forcing, (Newmark)]
            -> forcing -> streaming -> boundary conditions (in order) -> obstacle mask

``Stepper`` takes a declarative description of that recipe (``spec``) and advances it with one
fused pull-stream + moments + collision + forcing kernel per step (vsb_step), marker-parallel IB
kernels (vsb_ib_window_moments, vsb_ib_mdf) and an ordered wall-layer fix-up.

State convention (hard part 1).  The reference carries F_n (after streaming and
boundary conditions).  Internally the stepper carries S_n = collide(F_n), so each step is a pure
pull: S_{n+1} = collide(post(stream(S_n))).  ``set_f`` / ``get_f`` convert at the ends, so users
only ever see the reference's F.

spec keys (same dict the CPU oracle's ``oracle.recipes`` accepts)
    dim 2|3, shape, collision "bgk"|"mrt"|"kbc"|"reg", omega, forcing None|"edm"|"guo",
    g None | (gx, gy[, gz]) uniform | tensor (dim, *shape), mrt_op / mrt_fop optional Q x Q matrices,
    ib None | dict(markers (M, dim), ds scalar|(M,), kernel, n_iter, u_target None|(M, dim), window (origin, size)),
    post ordered list of (name, loc, kwargs) / ("mask", mask)
"""

import os
import ctypes as C

import numpy as np
import torch

from . import _api, _lib as L

_WRAPS = ("velocity", "pressure", "force_corrected")


def _parse_bc(name):
    for w in _WRAPS:
        if name.startswith(w + "_"):
            return name[len(w) + 1:], w
    return name, ""


TILE_CHUNK = 256          # markers per CTA of the tiled MDF kernel (kTiledChunk in csrc/vsb_ib.cu)
TILE_CELLS = int(os.environ.get("VSB_TILE_CELLS", "2304"))   # cells of its shared-memory box (kTileCells; tuning builds only)
TILE_COLUMN = int(os.environ.get("VSB_TILE_COLUMN", "4"))    # markers of one chunk share a TILE_COLUMN x TILE_COLUMN column of cells in (x, y)


def reachable_window_cells(markers, window_origin, window_size):
    """Ascending flat indices (int32) of the IB-window cells a 4-point stencil of some marker can ever touch while the
    window follows the body: the stencil nodes base - 1 .. base + 2 of every marker at rest, with the base allowed to
    drift by one cell either way (the window origin is an integer, the body's displacement is not).  Host logic,
    NumPy only.  For a finely meshed surface this is a thin shell of the window -- the only cells whose velocity has to
    be computed and whose force / work fields are ever written."""
    markers = np.asarray(markers, dtype=np.float64)
    dim = markers.shape[1]
    origin = np.floor(np.asarray(window_origin, dtype=np.float64)[:dim]).astype(np.int64)
    size = np.asarray(window_size, dtype=np.int64)[:dim]
    vol = np.zeros(tuple(int(k) for k in size), dtype=bool)
    if markers.shape[0]:
        base = np.clip(np.floor(markers - origin).astype(np.int64), 0, size - 1)
        vol[tuple(base[:, d] for d in range(dim))] = True
    for ax in range(dim):
        acc = np.zeros_like(vol)
        for sft in range(-2, 4):
            src = [slice(None)] * dim
            dst = [slice(None)] * dim
            if sft >= 0:
                src[ax], dst[ax] = slice(0, vol.shape[ax] - sft), slice(sft, vol.shape[ax])
            else:
                src[ax], dst[ax] = slice(-sft, vol.shape[ax]), slice(0, vol.shape[ax] + sft)
            acc[tuple(dst)] |= vol[tuple(src)]
        vol = acc
    return np.flatnonzero(vol).astype(np.int32)


def cut_marker_chunks(markers):
    """Storage order and chunk boundaries of a dense 3-D marker set for the tiled MDF kernel (host logic, NumPy only).

    Markers are sorted by (x column of 4 cells, y column of 4 cells, z); the sorted list is cut into chunks of at most
    TILE_CHUNK markers of ONE column whose z range keeps the chunk's bounding box -- column 4 + stencil 3 + 1 for a
    moving body = 8 cells in x and in y, z extent = span + 4 stencil cells + 1 margin -- within TILE_CELLS cells.
    Returns (perm, offsets): `markers[perm]` is the storage order, chunk c holds sorted markers
    [offsets[c], offsets[c + 1]) (int32, offsets[0] = 0, offsets[-1] = M)."""
    markers = np.asarray(markers, dtype=np.float32)
    col = np.floor(markers[:, :2] / float(TILE_COLUMN)).astype(np.int64)
    perm = np.lexsort((markers[:, 2], col[:, 1], col[:, 0]))
    srt = markers[perm]
    col = col[perm]
    key = col[:, 0] * (1 << 32) + col[:, 1]
    starts = np.flatnonzero(np.r_[True, key[1:] != key[:-1]])
    ends = np.r_[starts[1:], key.size]
    box_xy = TILE_COLUMN + 4
    z_span = TILE_CELLS // (box_xy * box_xy) - 5
    offsets = [0]
    for s_, e_ in zip(starts, ends):
        z = srt[s_:e_, 2]
        pos = 0
        while pos < e_ - s_:
            stop = min(pos + TILE_CHUNK, int(np.searchsorted(z, z[pos] + z_span, side="left")))
            stop = max(stop, pos + 1)
            offsets.append(s_ + stop)
            pos = stop
    return perm, np.asarray(offsets, dtype=np.int32)


class Stepper:
    def __init__(self, spec, device="cuda", rows=None, vec=0, body=None, dyn_mode="host", follow=1, use_graph=False,
                 fuse_ib=True, fuse_edges=True, overlap=True, buffers=None, ib_chain="auto", ib_shard=None,
                 chain_first=None, host_ode="poll"):
        """rows: (begin, end) range of the slowest axis that is physical domain (ghost layers outside; slab
        decomposition).  body: dict(m, k, c, added_mass, n_dof=2, d0, v0, a0) for a moving rigid body coupled by
        Newmark-beta; m, k, c scalars (dyn.py:44-46) or (n_dof, n_dof) matrices / length-n_dof diagonals
        (dyn.py:36-42), added_mass a scalar or one value per degree of freedom; in 2-D, n_dof=3 with
        rotation=True and center=(cx, cy) adds the rotation about the centre (dyn.py:84-154).
        dyn_mode "host" (reference-faithful, one tiny D2H/H2D per step) or "device"
        (vsb_body_newmark, graph-capturable).  follow: IB window rule for a moving body (1 trunc, 2 clip(floor)).
        ib_chain: how a small body's MDF iterations are chained in one launch -- "auto", "cta" (ONE CTA, work field in
        shared memory, spread as a gather over per-cell buckets: no floating-point atomics, bit-reproducible; 2-D,
        <= 512 markers), "barrier" (grid barriers, cooperative launch), "cluster" (one thread-block cluster, work fields in distributed shared memory; 2-D,
        <= 512 markers) or "launches" (one launch per iteration).
        chain_first: put a one-launch IB chain on the SMs BEFORE the bulk pass -- the bulk is enqueued behind it as a
        programmatic dependent launch and fills the rest of the device, the window band follows the chain on a second
        stream (None: on when the chain is one launch and the body's ODE is not on the host).
        host_ode: with dyn_mode="host", how the host learns that a step's force has arrived -- "poll" (the calling
        thread polls a page-locked mailbox inside vsb_run_host_ode: lowest latency, blocks until the steps are done) or
        "callback" (vsb_enqueue_host_ode: the Newmark update runs as a stream-ordered host function, step() returns
        at once).
        ib_shard: multidevice.IbShard -- this stepper is one slab of a decomposed run and shares the IB chain with the
        other ranks (spec['ib'] and the body then carry GLOBAL coordinates).
        fuse_ib / fuse_edges / overlap: use the single-kernel IB path, the single-kernel wall path and concurrent
        streams when the configuration allows (all three only change scheduling, not arithmetic per cell)."""
        L.lib()
        if not torch.cuda.is_available():
            raise L.VsbError("vivsim_b200.Stepper needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device(device)
        self.spec = spec
        self.dim = int(spec["dim"])
        self.shape = tuple(int(n) for n in spec["shape"])
        if self.dim not in (2, 3) or len(self.shape) != self.dim:
            raise ValueError("spec['dim'] must be 2 or 3 and match len(spec['shape'])")
        self.q = _api.Q[self.dim]
        self.rows = (0, self.shape[0]) if rows is None else (int(rows[0]), int(rows[1]))
        self.vec = int(vec)
        self.use_graph = bool(use_graph)
        self._keep = []
        if buffers is None:
            self._bufs = [torch.zeros((self.q,) + self.shape, device=self.device, dtype=torch.float32) for _ in range(2)]
        else:   # caller-owned ping-pong buffers (e.g. peer-mapped symmetric memory for the multi-GPU halo)
            self._bufs = list(buffers)
            for b in self._bufs:
                if tuple(b.shape) != (self.q,) + self.shape or b.dtype != torch.float32 or not b.is_cuda or not b.is_contiguous():
                    raise ValueError("buffers must be two contiguous fp32 CUDA tensors of shape (Q, *shape)")
        self._cur = 0
        self.n_steps = 0           # reference time steps taken since set_f / restore
        self._kind = None          # 'F' (reference state) or 'S' (post-collision state)
        self._tmp = None
        self._graph = None
        self.n_launch_per_step = 0
        self._parity = 0
        self.halo = None
        self._want = dict(fuse_ib=bool(fuse_ib), fuse_edges=bool(fuse_edges), overlap=bool(overlap))
        if ib_chain not in L.CHAIN:
            raise ValueError(f"ib_chain must be one of {sorted(L.CHAIN)}, got {ib_chain!r}")
        self._ib_chain = ib_chain
        self._chain_first_want = chain_first
        if host_ode not in ("poll", "callback"):
            raise ValueError(f"host_ode must be 'poll' or 'callback', got {host_ode!r}")
        self._host_ode = host_ode
        self._shard = ib_shard
        # extent the IB window and the body live in: the whole decomposed grid when the chain is shared
        self._gshape = tuple(ib_shard.slab.global_shape) if ib_shard is not None else self.shape
        self._xshift = (1 - ib_shard.slab.x0) if ib_shard is not None else 0
        self._side = None

        a = L.VsbStepArgs()
        a.grid = L.grid_of(self.shape)
        if spec["collision"] not in L.COLL:
            raise ValueError(f"unknown collision {spec['collision']!r}")
        a.collision = L.COLL[spec["collision"]]
        forcing = spec.get("forcing")
        if forcing not in L.FORCE:
            raise ValueError(f"unknown forcing {forcing!r}")
        a.forcing = L.FORCE[forcing]
        a.omega = float(spec["omega"])
        if spec["collision"] == "mrt":
            op = spec.get("mrt_op")
            self._op = L.host_matrix(op if op is not None else _api.mrt_operator(self.dim, a.omega), self.q)
            a.mrt_op_host = self._op.ctypes.data
            if forcing == "guo":
                fop = spec.get("mrt_fop")
                self._fop = L.host_matrix(fop if fop is not None else _api.mrt_operator(self.dim, a.omega, True), self.q)
                a.mrt_fop_host = self._fop.ctypes.data
        a.row_begin, a.row_end = self.rows
        a.vec = self.vec

        g = spec.get("g")
        self._g_field = None
        if g is not None and a.forcing:
            if getattr(g, "ndim", 1) == self.dim + 1:
                # a force field (dim, *shape): stored like the IB force window, cell-major with the components packed
                # (float2 in 2-D, float4 in 3-D), the window being the whole grid
                gt = L.dev(torch.as_tensor(np.asarray(g)).to(self.device) if not isinstance(g, torch.Tensor) else g, name="g")
                if tuple(gt.shape) != (self.dim,) + tuple(self.shape):
                    raise ValueError(f"g must have shape {(self.dim,) + tuple(self.shape)}, got {tuple(gt.shape)}")
                packed = torch.zeros(tuple(self.shape) + (2 if self.dim == 2 else 4,), device=self.device, dtype=torch.float32)
                packed[..., :self.dim] = gt.movedim(0, -1)
                self._g_field = packed.contiguous()
            else:
                gv = [float(x) for x in np.asarray(g, dtype=np.float64).reshape(-1)]
                if len(gv) != self.dim:
                    raise ValueError("uniform g needs one value per dimension")
                for d in range(self.dim):
                    a.g_uniform[d] = gv[d]

        # ---- immersed boundary
        self.ib = spec.get("ib")
        self.body = None
        self._body_dev = None
        self._hist = None
        self._perm = self._inv_perm = None
        if self.ib is not None:
            if not a.forcing:
                raise ValueError("an immersed boundary needs forcing='edm' or 'guo'")
            if self._g_field is not None:
                raise ValueError("a full force field and an IB window cannot be combined; add the field as uniform g")
            self._init_ib(a, body, dyn_mode, follow)
        elif self._g_field is not None:
            a.g_win = self._g_field.data_ptr()
            for d in range(self.dim):
                a.win_origin[d] = 0
                a.win_size[d] = self.shape[d]

        # ---- ordered post-streaming operations
        ops = []
        for item in spec.get("post", ()):
            if item[0] == "mask":
                m = item[1]
                m = torch.as_tensor(np.asarray(m), device=self.device) if not isinstance(m, torch.Tensor) else m.to(self.device)
                op, _ = _api.make_post_op(self.dim, self.shape, "mask", keep=self._keep, mask=m)
            else:
                kind, wrap = _parse_bc(item[0])
                kw = dict(item[2]) if len(item) > 2 else {}
                for k_, v_ in list(kw.items()):
                    if isinstance(v_, np.ndarray) and v_.ndim > 0:
                        kw[k_] = torch.as_tensor(v_, device=self.device, dtype=torch.float32)
                op, _ = _api.make_post_op(self.dim, self.shape, kind, item[1], wrap, keep=self._keep, **kw)
                self._check_window_clear_of(item[1])
            ops.append(op)
        self._post = (L.VsbPostOp * max(len(ops), 1))(*ops)
        a.n_post = len(ops)
        a.post = self._post
        self._args = a
        lib = L.lib()
        n_bc = sum(1 for o in ops if o.kind != L.BC["mask"])
        a.do_stream, a.do_collide = 1, 1
        a.f_in, a.f_out = self._bufs[0].data_ptr(), self._bufs[1].data_ptr()
        self.edge_fused = bool(n_bc) and self._want["fuse_edges"] and bool(lib.vsb_edge_fused_supported(C.byref(a)))
        self.ib_fused = (self.ib is not None and self._want["fuse_ib"] and self._shard is None
                         and bool(lib.vsb_ib_fused_supported(C.byref(self._mdf))))
        # concurrent branches need every kernel of a step to be independent of launch order
        self.overlap = (self._want["overlap"] and self.ib is not None and (n_bc == 0 or self.edge_fused))
        if self.overlap or (self.edge_fused and self._want["overlap"]):
            self._side = self._side_streams()
        self.n_launch_per_step = self._count_launches(n_bc)

    # ------------------------------------------------------------------ setup helpers
    def _side_streams(self):
        """IB and wall-layer streams.  The IB chain is short, serial and latency-bound while the bulk kernel keeps every
        SM full: a higher stream priority lets the chain's CTAs take freed slots first instead of queueing behind
        the bulk's remaining blocks (VSB_IB_PRIORITY=0 switches it off)."""
        prio = int(os.environ.get("VSB_IB_PRIORITY", "-1"))
        return [torch.cuda.Stream(device=self.device, priority=prio), torch.cuda.Stream(device=self.device)]

    def _init_ib(self, a, body, dyn_mode, follow):
        ib, dim, dev = self.ib, self.dim, self.device
        markers = np.asarray(ib["markers"], dtype=np.float32)
        if markers.ndim != 2 or markers.shape[1] != dim:
            raise ValueError(f"ib['markers'] must have shape (M, {dim})")
        origin, size = ib["window"]
        self.win_origin0 = tuple(float(o) for o in origin)
        self.win_size = tuple(int(n) for n in size)
        for d in range(dim):
            lo, n = int(np.floor(self.win_origin0[d])), self.win_size[d]
            if n < 4 or lo < 0 or lo + n > self._gshape[d]:
                raise ValueError(f"IB window axis {d}: [{lo}, {lo + n}) must lie inside the grid [0, {self._gshape[d]})")
        if body is None:
            rel = markers - np.floor(np.asarray(self.win_origin0, dtype=np.float32))
            if (np.floor(rel).min(axis=0) < 1).any() or (np.floor(rel).max(axis=0) + 2 >= np.asarray(self.win_size)).any():
                raise ValueError("every marker's 4-point stencil must lie inside the IB window (the reference leaves "
                                 "out-of-range stencil indices undefined)")
        self.n_markers = markers.shape[0]
        self.n_iter = int(ib.get("n_iter", 5))
        # dense marker sets (many stencil points per window cell, e.g. a finely meshed 3-D surface) interpolate a
        # precomputed window velocity; sparse ones take it from the streamed populations at their stencil points
        wcells = int(np.prod(self.win_size))
        self._use_uwin = self.n_markers * 4 ** dim > 2 * wcells or self._shard is not None
        # The tiled MDF kernel (dense 3-D bodies) gives each CTA 256 consecutive markers and privatises their part of
        # the window in shared memory, so consecutive markers must be close in space: store them sorted by
        # (x column of 4 cells, y column of 4 cells, z).  Outputs are handed back in the caller's order.
        self._perm = self._inv_perm = None
        self._chunk_offsets = None
        if self._shard is not None:
            # shared chain: the plan fixes the storage order (rank shares are contiguous) and the chunks
            plan = self._shard.plan
            perm = plan["perm"]
            inv = np.empty_like(perm)
            inv[perm] = np.arange(perm.size)
            markers = np.ascontiguousarray(markers[perm])
            self._perm = torch.as_tensor(perm, device=dev)
            self._inv_perm = torch.as_tensor(inv, device=dev)
            if plan["chunk_offsets"] is not None:
                self._chunk_offsets = torch.as_tensor(plan["chunk_offsets"], device=dev)
        elif dim == 3 and self._use_uwin and self.n_markers > 480 and ib.get("sort_markers", True):
            perm, offsets = cut_marker_chunks(markers)
            inv = np.empty_like(perm)
            inv[perm] = np.arange(perm.size)
            markers = np.ascontiguousarray(markers[perm])
            self._perm = torch.as_tensor(perm, device=dev)
            self._inv_perm = torch.as_tensor(inv, device=dev)
            self._chunk_offsets = torch.as_tensor(offsets, device=dev)
        self._markers = torch.as_tensor(markers, device=dev)
        # force field [0] and per-iteration work fields [1:], double-buffered by step parity: each step clears the
        # set the next step will accumulate into (see vsb_ib_mdf), so there is no memset on the step path
        # layout: cell-major with the components packed per cell (float2 in 2-D, float4 in 3-D)
        nc = 2 if dim == 2 else 4
        if self._shard is not None:
            self._ib_buf = self._shard.fields           # (2, n_iter + 1, *window, nc), peer-mapped; slot n_iter = velocity
            if tuple(self._ib_buf.shape) != (2, self.n_iter + 1) + self.win_size + (nc,):
                raise ValueError("ib_shard was built for another window / n_iter")
            self._u_win = None
        else:
            self._ib_buf = torch.zeros((2, self.n_iter) + self.win_size + (nc,), device=dev)
            self._u_win = torch.zeros(self.win_size + (nc,), device=dev) if self._use_uwin else None
        self._g_win = self._ib_buf[0, 0] if self._shard is None else self._shard.force_field
        self._marker_u = torch.zeros((self.n_markers, dim), device=dev)
        self._marker_force = torch.zeros((self.n_markers, dim), device=dev)   # +F; reaction on the body is -F
        tgt = ib.get("u_target")
        if tgt is not None and self._perm is not None:
            tgt = np.asarray(tgt, dtype=np.float32)[self._perm.cpu().numpy()]
        self._u_target = None if tgt is None else torch.as_tensor(np.asarray(tgt, dtype=np.float32), device=dev)
        ds = ib["ds"]
        self._ds = None
        m = L.VsbMdfArgs()
        m.dim, m.delta_kind, m.n_iter = dim, L.DELTA[ib.get("kernel", "peskin4")], self.n_iter
        m.n_markers = self.n_markers
        if np.ndim(ds) > 0:
            ds = np.asarray(ds, dtype=np.float32)
            if self._perm is not None and ds.shape == (self.n_markers,):
                ds = ds[self._perm.cpu().numpy()]
            self._ds = torch.as_tensor(np.asarray(ds, dtype=np.float32), device=dev)
            if self._ds.shape != (self.n_markers,):
                raise ValueError("ib['ds'] must be a scalar or have shape (M,)")
            m.ds_ptr = self._ds.data_ptr()
        else:
            m.ds_value = float(ds)
        for d in range(dim):
            m.win_origin0[d] = int(np.floor(self.win_origin0[d]))
            m.win_size[d] = self.win_size[d]
            a.win_origin[d] = m.win_origin0[d] + (self._xshift if d == 0 else 0)   # the fluid kernels index the local slab
            a.win_size[d] = self.win_size[d]
        a.win_shift[0] = self._xshift
        if self._chunk_offsets is not None:
            m.chunk_offsets = self._chunk_offsets.data_ptr()
            m.n_chunks = self._chunk_offsets.numel() - 1
        m.markers0 = self._markers.data_ptr()
        m.u_target = self._u_target.data_ptr() if self._u_target is not None else None
        self._mdf_barrier = torch.zeros(2, dtype=torch.int64, device=dev)
        m.barrier = self._mdf_barrier.data_ptr()
        lanes = 16 if dim == 2 else 32
        chain = self._ib_chain
        if chain == "cluster" and (dim != 2 or self.n_markers > 512):
            raise ValueError("ib_chain='cluster' is for 2-D bodies of at most 512 markers")
        if chain == "cluster":
            # the cluster kernel keeps, per marker, the markers whose 4 x 4 stencil can overlap its own (within 3 cells
            # per axis, whatever the rigid motion): at most 96.  The marker set is rigid, so check once.
            d2 = ((markers[:, None, :].astype(np.float64) - markers[None, :, :]) ** 2).sum(axis=2)
            near = d2 < (4 * 2 ** 0.5 + 1.5) ** 2            # |base difference| <= 3 per axis => distance < 4 sqrt(2) + 1
            stride = int(near.sum(axis=1).max())
            if stride > 48:
                raise ValueError("ib_chain='cluster': more than 48 markers within reach of one marker's stencil")
            stride = (stride + 3) // 4 * 4
            nbr = np.full((stride, 512), 0xFFFF, dtype=np.uint16)      # neighbour-major, one column per marker
            for i in range(self.n_markers):
                js = np.flatnonzero(near[i])
                nbr[:js.size, i] = js
            self._nbr = torch.as_tensor(nbr.view(np.int16), device=dev)
            m.nbr_list, m.nbr_stride = self._nbr.data_ptr(), stride
        # Dense body in a window that follows it: only a thin shell of the window is ever within reach of a stencil.
        # Its cell list lets the window-velocity kernel and the per-step clearing skip the rest (C5: 7.1 M -> 0.9 M cells).
        self._reach = None
        rigid_in_window = body is None or (int(follow) in (1, 2) and not bool(body.get("rotation", False)))
        if self._use_uwin and self._shard is None and rigid_in_window and ib.get("reach_list", True):
            cells = reachable_window_cells(markers, self.win_origin0, self.win_size)
            if cells.size < 0.5 * wcells:
                self._reach = torch.as_tensor(cells, device=dev)
                m.reach_cells, m.n_reach_cells = self._reach.data_ptr(), int(cells.size)
        m.chain_mode = L.CHAIN[chain]
        self._mdf_one_launch = ((self.n_markers * lanes + 127) // 128 <= 120 and self._ib_chain != "launches"
                                and self._shard is None)
        m.marker_u = self._marker_u.data_ptr()
        m.marker_force = self._marker_force.data_ptr()
        a.g_win = self._g_win.data_ptr()
        self.follow = 0
        self._bparams = None
        if body is not None:
            self.body = dict(body)
            self.dyn_mode = dyn_mode
            if dyn_mode not in ("host", "device"):
                raise ValueError("dyn_mode must be 'host' or 'device'")
            self.follow = int(follow)
            self.n_dof = int(body.get("n_dof", 2))
            self.rotation = bool(body.get("rotation", False))
            if self.rotation and (dim != 2 or self.n_dof != 3 or "center" not in body):
                # the reference's rigid-body kinematics with rotation are 2-D (dyn.py:84-154)
                raise ValueError("body['rotation'] needs a 2-D body with n_dof=3 and center=(cx, cy)")
            if not 1 <= self.n_dof <= (3 if self.rotation else dim):
                raise ValueError(f"body['n_dof'] must be 1..{dim} translational degrees of freedom in {dim}-D (3 with "
                                 f"rotation=True in 2-D), got {self.n_dof}")
            bp = L.VsbBodyParams()
            bp.n_dof, bp.follow = (self.n_dof if dyn_mode == "device" else 0), self.follow
            mats = {key: np.asarray(body[key], dtype=np.float64) for key in ("m", "k", "c")}
            added = np.asarray(body.get("added_mass", 0.0), dtype=np.float64)
            if any(v.ndim > 0 for v in mats.values()) or added.ndim > 0:
                bp.matrix_form = 1
                n = self.n_dof
                for key, field in (("m", bp.mat_m), ("k", bp.mat_k), ("c", bp.mat_c)):
                    v = mats[key]
                    full = v * np.eye(n) if v.ndim == 0 else (np.diag(v) if v.ndim == 1 else v)
                    if full.shape != (n, n):
                        raise ValueError(f"body[{key!r}] must be a scalar, {n} diagonal entries or an ({n}, {n}) matrix")
                    for i in range(n):
                        for j in range(n):
                            field[3 * i + j] = float(full[i, j])
                eff = np.array([[bp.mat_m[3 * i + j] + 0.5 * bp.mat_c[3 * i + j] + 0.25 * bp.mat_k[3 * i + j]
                                 for j in range(n)] for i in range(n)])
                if abs(np.linalg.det(eff)) == 0.0:
                    raise ValueError("m + c/2 + k/4 is singular")
                av = np.broadcast_to(added, (n,))
                for i in range(n):
                    bp.added_mass_v[i] = float(av[i])
            else:
                bp.m, bp.k, bp.c, bp.added_mass = float(mats["m"]), float(mats["k"]), float(mats["c"]), float(added)
            if self.rotation:
                bp.rotation = 1
                bp.center[0], bp.center[1] = (float(x) for x in body["center"])
                m.rotation = 1
                m.center[0], m.center[1] = bp.center[0], bp.center[1]
            if self._shard is not None and dyn_mode != "device":
                raise ValueError("a body whose IB chain is shared among ranks needs dyn_mode='device' (every rank "
                                 "advances an identical replica of the rigid-body state)")
            for d in range(3):
                bp.origin0[d] = self.win_origin0[d] if d < dim else 0.0
                bp.grid_size[d] = self._gshape[d] if d < dim else 1
                bp.win_size[d] = self.win_size[d] if d < dim else 1
            self._bparams = bp
            # optional per-step record of (d, h) (what the reference's update_chunk scan returns): a ring written by
            # the body update itself -- device memory for the device ODE, page-locked host memory for the host ODE
            self._hist_cap = int(body.get("history", 0))
            self._hist = None
            if self._hist_cap > 0:
                if dyn_mode == "device":
                    self._hist = torch.zeros((self._hist_cap, 6), device=dev, dtype=torch.float32)
                else:
                    self._hist = torch.zeros((self._hist_cap, 6), dtype=torch.float32).pin_memory()
                bp.history, bp.history_capacity = self._hist.data_ptr(), self._hist_cap
            hp = L.VsbBodyParams.from_buffer_copy(bp)      # host-ODE variant: same numbers, n_dof always set
            hp.n_dof = self.n_dof
            self._hparams = hp
            self._body_dev = torch.zeros(L.BODY_BYTES // 4, device=dev, dtype=torch.float32)
            self._body_pin = torch.zeros(L.BODY_BYTES // 4, dtype=torch.float32).pin_memory()
            st = self._body_pin.numpy()
            st[0:self.n_dof] = np.asarray(body.get("d0", np.zeros(self.n_dof)), dtype=np.float32)
            st[3:3 + self.n_dof] = np.asarray(body.get("v0", np.zeros(self.n_dof)), dtype=np.float32)
            st[6:6 + self.n_dof] = np.asarray(body.get("a0", np.zeros(self.n_dof)), dtype=np.float32)
            org = self._origin_for(st[0:3])
            ints = st.view(np.int32)
            ints[15:18] = org
            ints[18:21] = org
            self._body_dev.copy_(self._body_pin, non_blocking=True)
            if dyn_mode == "host":
                # mailbox in page-locked host memory: the last CTA of the MDF chain posts the total force there and
                # the host polls it (vsb_run_host_ode) -- no device->host copy, no stream synchronisation per step
                self._mail = torch.zeros(8, dtype=torch.int32).pin_memory()
                self._mail[4] = 1                      # VsbHostMail.next
                m.host_mail = self._mail.data_ptr()
            m.body = self._body_dev.data_ptr()
            a.body = self._body_dev.data_ptr()
        self._mdf = m

    def _origin_for(self, d):
        """Integer window origin for displacement d (same fp32 rule as origin_rule in csrc/vsb_step.cuh)."""
        out = np.zeros(3, dtype=np.int32)
        for k in range(self.dim):
            shifted = np.float32(self.win_origin0[k]) + (np.float32(d[k]) if self.follow else np.float32(0))
            o = int(np.floor(shifted)) if self.follow == 2 else int(shifted)   # int(): truncation like astype(int32)
            if self.follow:             # inside the grid, like the start of a lax.dynamic_slice
                o = max(min(o, self._gshape[k] - self.win_size[k]), 0)
            out[k] = o
        return out

    def _check_window_clear_of(self, loc):
        if self.ib is None:
            return
        axis, low = L.LOC[loc] // 2, L.LOC[loc] % 2 == 0
        lo = int(np.floor(self.win_origin0[axis]))
        hi = lo + self.win_size[axis]
        rb, re = (self.rows if axis == 0 else (0, self.shape[axis]))
        if axis == 0 and self._shard is not None:      # global coordinates: only the ends of the whole grid are walls
            rb, re = 0, self._gshape[0]
        if (low and lo <= rb) or (not low and hi >= re):
            raise ValueError(f"the IB window touches the '{loc}' wall layer, which carries a boundary operation")

    def _count_launches(self, n_bc):
        n = 1
        if n_bc:
            n += 0 if self.edge_fused else 2 + self._args.n_post
        if self.ib is not None:
            if self.overlap:
                n += 1                                    # second launch of the fused kernel (window x-range)
            n += 1 if (self.ib_fused or self._mdf_one_launch) else self.n_iter
            n += 1 if (self._use_uwin and not self.ib_fused) else 0
            n += (self.n_iter + 3) if self._shard is not None else 0      # flag barriers, force push and reduce of the shared chain
        return n

    def attach_halo(self, halo, fused=None):
        """Multi-GPU slabs: `halo` (multidevice.PeerHalo) exchanges ghost layers inside every collide pass.  When all
        face operations sit on x faces (or there are none) the exchange is pipelined: interior rows start at once,
        while a second stream waits for the neighbours, updates the two edge rows and sends them on.  fused (default:
        on, VSB_HALO_FUSED=0 turns it off): the edge-row launch does the wait and the send itself (VsbStepArgs.halo)
        instead of a vsb_halo_wait before and a vsb_halo_send after it."""
        self.halo = halo
        locs = [self._post[i].loc for i in range(self._args.n_post) if self._post[i].kind != L.BC["mask"]]
        x_only = all(loc in (L.LOC["left"], L.LOC["right"]) for loc in locs)
        self.halo_pipelined = (self.rows[1] - self.rows[0] >= 4) and (not locs or (self.edge_fused and x_only))
        if self._side is None:
            self._side = self._side_streams()
        self._halo_stream = torch.cuda.Stream(device=self.device)
        self._chain_done = torch.cuda.Event()
        if fused is None:
            fused = os.environ.get("VSB_HALO_FUSED", "1") != "0"
        self.halo_fused = bool(fused and self.halo_pipelined and getattr(halo, "args", None) is not None)
        self.n_launch_per_step += (1 if self.halo_fused else 3) if self.halo_pipelined else 1

    # ------------------------------------------------------------------ state access
    def set_f(self, f):
        """Load the reference-convention state F (post-streaming, post-boundary populations)."""
        f = torch.as_tensor(np.asarray(f), device=self.device) if not isinstance(f, torch.Tensor) else f.to(self.device)
        if tuple(f.shape) != (self.q,) + self.shape:
            raise ValueError(f"f must have shape {(self.q,) + self.shape}, got {tuple(f.shape)}")
        self._bufs[self._cur].copy_(f.to(torch.float32))
        self._kind = "F"
        self.n_steps = 0
        return self

    def _drop_graph(self):
        """A captured graph replays from the (buffer, parity) pair it was recorded with.  After restore() the live pair
        may be the other one; rather than rely on one realigning step, re-capture from the restored state."""
        self._graph = None

    def get_f(self):
        """Return F_n, the state the reference carries after n steps (a new tensor)."""
        self._require_state()
        if self._kind == "F":
            return self._bufs[self._cur].clone()
        out = torch.empty_like(self._bufs[0])
        self._launch(self._bufs[self._cur], out, do_stream=1, do_collide=0)
        return out

    @property
    def marker_force(self):
        """+F on every marker, shape (M, dim), in the order the markers were given (the reaction on the body is -F)."""
        if self._inv_perm is None:
            return self._marker_force
        return self._marker_force.index_select(0, self._inv_perm)

    @property
    def state(self):
        """The internal buffer (S_n after the first step, F_0 before).  For halo exchange and tests."""
        return self._bufs[self._cur]

    def body_state(self):
        """(d, v, a, h) of the rigid body as NumPy arrays (synchronises)."""
        st = self._body_dev.cpu().numpy()
        n = self.n_dof
        return st[0:n].copy(), st[3:3 + n].copy(), st[6:6 + n].copy(), st[9:9 + n].copy()

    def body_steps(self):
        """Number of body updates performed so far (synchronises)."""
        return int(self._body_dev.cpu().numpy().view(np.int32)[22])

    def body_history(self, n=None):
        """(d, h) of the last n steps, each of shape (n, n_dof), oldest first -- the per-step record the reference's
        update_chunk returns (examples/2d/vortex_induced_vibration.py:150-157).  Needs body['history'] = capacity."""
        if self._body_dev is None or self._hist is None:
            raise L.VsbError("no history: pass body=dict(..., history=capacity)")
        done = self.body_steps()          # also synchronises the device-side writes
        if self._hist.is_cuda:
            ring = self._hist.cpu().numpy()
        else:
            torch.cuda.synchronize(self.device)
            ring = self._hist.numpy().copy()
        n = min(done, self._hist_cap) if n is None else int(n)
        if n > min(done, self._hist_cap):
            raise ValueError(f"only {min(done, self._hist_cap)} steps are recorded (capacity {self._hist_cap})")
        rows = ring[[(done - n + i) % self._hist_cap for i in range(n)]]
        return rows[:, 0:self.n_dof].copy(), rows[:, 3:3 + self.n_dof].copy()

    # ------------------------------------------------------------------ dump / restore
    def checkpoint(self):
        """Everything needed to resume bit-identically, as a dict of NumPy arrays: the raw population buffer and its
        convention, the step parity, the rigid-body state and the marker arrays.  (The reference's examples carry
        (f, d, v, a) between chunks, examples/2d/vortex_induced_vibration.py:150-157,196-205.)"""
        self._require_state()
        torch.cuda.synchronize(self.device)
        ck = {"populations": self._bufs[self._cur].cpu().numpy(), "kind": np.array(self._kind),
              "parity": np.array(self._parity, dtype=np.int32), "n_steps": np.array(self.n_steps, dtype=np.int64),
              "shape": np.array(self.shape, dtype=np.int64)}
        if self.ib is not None:
            ck["marker_force"] = self._marker_force.cpu().numpy()
            ck["marker_u"] = self._marker_u.cpu().numpy()
        if self._body_dev is not None:
            ck["body"] = self._body_dev.cpu().numpy().view(np.uint8).copy()
            if self._hist is not None:
                ck["history"] = self._hist.cpu().numpy()
        return ck

    def restore(self, ck):
        """Load a checkpoint() into a Stepper built from the same spec."""
        if tuple(int(n) for n in ck["shape"]) != self.shape or ck["populations"].shape[0] != self.q:
            raise ValueError(f"checkpoint is for a {tuple(ck['shape'])} grid, this stepper is {self.shape}")
        if (self._body_dev is not None) != ("body" in ck) or (self.ib is not None) != ("marker_force" in ck):
            raise ValueError("checkpoint and stepper disagree about the immersed body")
        self._drop_graph()
        self._bufs[self._cur].copy_(torch.as_tensor(np.ascontiguousarray(ck["populations"])))
        self._kind = str(ck["kind"])
        self.n_steps = int(ck["n_steps"])
        if self.ib is not None:
            self._parity = int(ck["parity"])
            self._marker_force.copy_(torch.as_tensor(ck["marker_force"]))
            self._marker_u.copy_(torch.as_tensor(ck["marker_u"]))
            # between steps the set the next step accumulates into is zero and the other one is cleared by the next
            # step before it is used again, so starting from all-zero work fields is equivalent
            self._ib_buf.zero_()
        if self._body_dev is not None:
            raw = np.ascontiguousarray(ck["body"]).view(np.float32)
            if raw.size != self._body_dev.numel():
                raise ValueError("checkpoint holds a body state of another ABI version")
            self._body_dev.copy_(torch.as_tensor(raw))
            self._body_pin.copy_(torch.as_tensor(raw))          # the host-ODE path keeps its master copy here
            if self._hist is not None and "history" in ck and ck["history"].shape == tuple(self._hist.shape):
                self._hist.copy_(torch.as_tensor(ck["history"]))
        torch.cuda.synchronize(self.device)
        return self

    def save(self, path):
        """checkpoint() to an .npz file."""
        np.savez(path, **self.checkpoint())

    def load(self, path):
        """restore() from a file written by save()."""
        with np.load(path) as z:
            return self.restore({k: z[k] for k in z.files})

    def macroscopic(self):
        """(rho, u) of the reference-convention state F_n (two passes: epilogue, moments)."""
        return _api.get_macroscopic(self.dim, self.get_f())

    def window_origin(self):
        """Integer IB-window origin that the next step will use."""
        if self._body_dev is None:
            return tuple(int(np.floor(o)) for o in self.win_origin0)
        ints = self._body_dev.cpu().numpy().view(np.int32)
        return tuple(int(x) for x in ints[15 + 3 * self._parity:15 + 3 * self._parity + self.dim])

    def _require_state(self):
        if self._kind is None:
            raise L.VsbError("no state loaded: call set_f(f) first")

    # ------------------------------------------------------------------ stepping
    def _launch(self, src, dst, do_stream, do_collide):
        """Enqueue one pass: (IB force) + fused kernel + wall layers.  Branches that do not depend on each other go
        to side streams when `overlap` is on: the bulk of the grid does not wait for the immersed boundary."""
        a = self._args
        a.f_in, a.f_out = src.data_ptr(), dst.data_ptr()
        a.do_stream, a.do_collide = do_stream, do_collide
        a.parity = self._parity
        if self.ib is not None and do_collide:     # force field of this step (double-buffered by parity)
            self._g_win = self._ib_buf[self._parity, 0] if self._shard is None else self._shard.force_field
            a.g_win = self._g_win.data_ptr()
        lib = L.lib()
        has_ops = a.n_post > 0 and do_stream
        with_ib = self.ib is not None and do_collide
        a.edges = 2 if (has_ops and self.edge_fused) else 0     # wall layers ride in the fused kernel's launch
        a.band = 0
        main = torch.cuda.current_stream()
        st_main = C.c_void_p(main.cuda_stream)
        ref = C.byref(a)
        halo = self.halo if do_collide else None
        dst_index = 1 - self._cur
        pipelined = halo is not None and self.halo_pipelined
        halo_waited = False
        rb, re = self.rows
        if pipelined:
            s_halo = self._halo_stream
            st_halo = C.c_void_p(s_halo.cuda_stream)
            s_halo.wait_stream(main)
            a.sub_begin, a.sub_end = rb + 1, re - 1            # interior rows: no ghost-layer dependency
        if (self.overlap and with_ib and not pipelined and self.body is not None and self.dyn_mode == "host"
                and not self.ib_fused and (not has_ops or self.edge_fused)):
            self._host_ode_step(main)                          # the whole pass in one C call
        elif not (self.overlap and with_ib):
            if with_ib:
                if pipelined and self._shard is not None:      # the shared chain reads the ghost rows: neighbours first
                    halo.wait(st_halo)
                    main.wait_stream(s_halo)
                    halo_waited = True
                self._ib_part(st_main)
            L.check(lib.vsb_step(ref, st_main))
        else:
            s_ib, s_edge = self._side
            st_ib = C.c_void_p(s_ib.cuda_stream)
            s_ib.wait_stream(main)
            if pipelined and self._shard is not None:
                # the shared chain reads this slab's ghost rows (window velocity of the edge rows): neighbours first
                halo.wait(st_halo)
                s_ib.wait_stream(s_halo)
            host_body = self.body is not None and self.dyn_mode == "host"
            if host_body:
                # the host ODE synchronises the IB stream: get the bulk going first so it runs meanwhile
                a.band = 1
                L.check(lib.vsb_step(ref, st_main))
            self._ib_part(st_ib)                               # IB chain (its few CTAs should not queue behind
            if pipelined and self._shard is not None:          # the bulk), then the window's x-range
                self._chain_done.record(s_ib)
            if self._chain_first() and not host_body and not pipelined:
                # The chain is ONE small launch: let it take its SMs first.  The bulk follows on the same stream as a
                # programmatic dependent launch (it starts once every CTA of the chain is resident, not when the
                # chain ends) and the band waits for the chain on the main stream.  Enqueued the other way round, the
                # bulk's ~1200 CTAs occupy every SM first and the chain's CTAs trickle in as slots free up.
                self._chain_ev.record(s_ib)
                a.band, a.early_launch = 1, 1
                L.check(lib.vsb_step(ref, st_ib))
                a.early_launch = 0
                main.wait_event(self._chain_ev)
                a.band = 2
                L.check(lib.vsb_step(ref, st_main))
                main.wait_stream(s_ib)
                a.band = 0
                if with_ib:
                    self._parity ^= 1
                return
            a.band = 2
            L.check(lib.vsb_step(ref, st_ib))
            if not host_body:
                a.band = 1                                     # everything but the window's x-range
                if self._shard is not None:
                    # shared chain: its first kernel (window velocity; a no-op on ranks whose slab holds no window cell)
                    # gets the memory system to itself -- every other rank waits for it at the first flag barrier
                    main.wait_event(self._shard.ev_window_done)
                L.check(lib.vsb_step(ref, st_main))
            main.wait_stream(s_ib)
            a.band = 0
        if pipelined:
            # second stream: neighbours' ghost layers -> the two edge rows (and the x walls) -> send them on
            fused = self.halo_fused
            hmode = 2                                          # the edge-row launch sends; 3: it waits first as well
            if self._shard is not None and with_ib and self.overlap:
                s_halo.wait_event(self._chain_done)            # a window on a cut: the edge rows read its force field
            elif halo_waited:
                s_halo.wait_stream(main)                       # chain and interior rows ran on `main`
            elif fused:
                hmode = 3
            else:
                halo.wait(st_halo)
            a.band = 0
            a.sub_begin, a.sub_end, a.edge_rows_only = 0, 0, 1
            if fused:
                a.halo, a.halo_mode = C.addressof(halo.args[dst_index]), hmode
            L.check(lib.vsb_step(ref, st_halo))
            a.edge_rows_only = 0
            if fused:
                a.halo, a.halo_mode = None, 0
            else:
                halo.send(dst_index, st_halo)
            main.wait_stream(s_halo)
        elif halo is not None:
            halo.push(dst_index, st_main)
        if with_ib:
            self._parity ^= 1

    def _chain_first(self):
        if getattr(self, "_chain_first_on", None) is None:
            want = self._chain_first_want
            if want is None:
                want = os.environ.get("VSB_CHAIN_FIRST", "0") == "1"
            self._chain_first_on = bool(want and self.ib is not None and self._mdf_one_launch and not self._use_uwin
                                        and not self.ib_fused and self._shard is None and self.halo is None)
            self._chain_ev = torch.cuda.Event()
        return self._chain_first_on

    def _host_ode_step(self, main):
        """One pass with the rigid-body ODE on the host through vsb_step_host_ode (fork / join in C)."""
        a, m = self._args, self._mdf
        if getattr(self, "_plan", None) is None:
            self._plan_events = [torch.cuda.Event() for _ in range(3)]
            for ev in self._plan_events:
                ev.record(main)                                # materialise the cudaEvent_t handles
            self._plan = L.VsbHostPlan()
            self._plan.ib, self._plan.edge = self._side[0].cuda_stream, self._side[1].cuda_stream
            self._plan.ev_fork, self._plan.ev_ib, self._plan.ev_edge = (ev.cuda_event for ev in self._plan_events)
        self._plan.main = main.cuda_stream
        par = self._parity
        m.parity = par
        buf = self._ib_buf
        m.g_win, m.g_win_next = buf[par, 0].data_ptr(), buf[par ^ 1, 0].data_ptr()
        if self.n_iter > 1:
            m.scratch, m.scratch_next = buf[par, 1].data_ptr(), buf[par ^ 1, 1].data_ptr()
        m.u_win = None
        L.check(L.lib().vsb_step_host_ode(C.byref(a), C.byref(m), C.byref(self._hparams),
                                          C.c_void_p(self._body_pin.data_ptr()), C.byref(self._plan)))

    def _host_ode_eligible(self):
        a = self._args
        return (self.overlap and self.ib is not None and self.body is not None and self.dyn_mode == "host"
                and self.halo is None and not self.ib_fused and (a.n_post == 0 or self.edge_fused))

    def _host_ode_prepare(self, main):
        """Argument blocks of the first step of a vsb_run_host_ode[_multi] call whose bulk runs on stream `main`."""
        a, m = self._args, self._mdf
        src, dst = self._bufs[self._cur], self._bufs[1 - self._cur]
        a.f_in, a.f_out = src.data_ptr(), dst.data_ptr()
        a.do_stream, a.do_collide = 1, 1
        a.edges = 2 if (a.n_post > 0 and self.edge_fused) else 0
        a.band = 0
        a.sub_begin, a.sub_end, a.edge_rows_only = 0, 0, 0
        if getattr(self, "_plan", None) is None:
            self._plan_events = [torch.cuda.Event() for _ in range(3)]
            for ev in self._plan_events:
                ev.record(main)                                # materialise the cudaEvent_t handles
            self._plan = L.VsbHostPlan()
            self._plan.ib, self._plan.edge = self._side[0].cuda_stream, self._side[1].cuda_stream
            self._plan.ev_fork, self._plan.ev_ib, self._plan.ev_edge = (ev.cuda_event for ev in self._plan_events)
        self._plan.main = main.cuda_stream
        par = self._parity
        a.parity = m.parity = par
        buf = self._ib_buf
        m.g_win, m.g_win_next = buf[par, 0].data_ptr(), buf[par ^ 1, 0].data_ptr()
        if self.n_iter > 1:
            m.scratch, m.scratch_next = buf[par, 1].data_ptr(), buf[par ^ 1, 1].data_ptr()
        m.u_win = None
        a.g_win = m.g_win

    def _host_ode_finish(self, n):
        """Python-side bookkeeping after n steps taken inside vsb_run_host_ode[_multi]."""
        self._cur = (self._cur + n) % 2
        self._parity = (self._parity + n) % 2
        self._g_win = self._ib_buf[self._parity, 0]

    def _run_host_ode(self, n):
        """n whole steps with the rigid-body ODE on the host in one C call (vsb_run_host_ode): the host loop, the
        mailbox polling and the Newmark update all happen in C; Python only flips its buffer / parity bookkeeping."""
        self._host_ode_prepare(torch.cuda.current_stream())
        fn = L.lib().vsb_enqueue_host_ode if self._host_ode == "callback" else L.lib().vsb_run_host_ode
        L.check(fn(C.byref(self._args), C.byref(self._mdf), C.byref(self._hparams),
                   C.c_void_p(self._body_pin.data_ptr()), C.byref(self._plan), int(n)))
        self._host_ode_finish(n)

    def _ib_part(self, st):
        """Immersed-boundary force of this pass on stream `st`, then the body update."""
        lib, a, m = L.lib(), self._args, self._mdf
        par = self._parity
        m.parity = par
        buf = self._ib_buf
        m.g_win, m.g_win_next = buf[par, 0].data_ptr(), buf[par ^ 1, 0].data_ptr()
        if self.n_iter > 1:
            m.scratch, m.scratch_next = buf[par, 1].data_ptr(), buf[par ^ 1, 1].data_ptr()
        bp = C.byref(self._bparams) if self._bparams is not None else None
        if self._shard is not None:        # chain shared with the other ranks (window fields in peer memory)
            L.check(lib.vsb_ibshard_chain(C.byref(a), C.byref(m), C.byref(self._shard.args), bp, st))
            return
        if self.ib_fused:
            L.check(lib.vsb_ib_fused(C.byref(a), C.byref(m), bp, st))
        else:
            m.u_win = None
            if self._use_uwin:
                if self._reach is not None:
                    L.check(lib.vsb_ib_window_moments_cells(C.byref(a), L.ptr(self._u_win), L.ptr(self._reach),
                                                            C.c_int64(self._reach.numel()), st))
                else:
                    L.check(lib.vsb_ib_window_moments(C.byref(a), L.ptr(self._u_win), st))
                m.u_win = self._u_win.data_ptr()
            L.check(lib.vsb_ib_mdf(C.byref(a), C.byref(m), bp, st))
        if self.body is not None and self.dyn_mode == "host":
            # rigid-body ODE on the host (north_star): 92 B device -> host, Newmark-beta on the CPU, 92 B back
            L.check(lib.vsb_body_newmark_host(C.c_void_p(self._body_dev.data_ptr()), C.c_void_p(self._body_pin.data_ptr()),
                                              C.byref(self._hparams), par, st))

    def _advance(self):
        src, dst = self._bufs[self._cur], self._bufs[1 - self._cur]
        self._launch(src, dst, do_stream=1, do_collide=1)
        self._cur = 1 - self._cur

    def advance_raw(self, n=1):
        """Enqueue n fused steps on the current stream with no graph handling (for callers that capture
        their own CUDA graph).  Needs the internal post-collision state, i.e. at least one step() before."""
        if self._kind != "S":
            raise L.VsbError("advance_raw needs the internal state: call set_f(f) and step(1) first")
        for _ in range(int(n)):
            self._advance()
        self.n_steps += int(n)
        return self

    def step(self, n=1):
        """Advance n reference time steps."""
        self._require_state()
        n = int(n)
        if n <= 0:
            return self
        self.n_steps += n
        if self._kind == "F":   # prologue: S_0 = collide(F_0)
            src, dst = self._bufs[self._cur], self._bufs[1 - self._cur]
            self._launch(src, dst, do_stream=0, do_collide=1)
            self._cur = 1 - self._cur
            self._kind = "S"
            n -= 1
        if n > 0 and self._host_ode_eligible():
            self._run_host_ode(n)
            return self
        graphable = self.use_graph and not (self.body is not None and self.dyn_mode == "host")
        if graphable and n >= 2:
            if self._graph is None:
                self._capture()
            if (self._graph_cur, self._graph_parity) != (self._cur, self._parity):   # recorded from the other buffer / parity
                self._advance()
                n -= 1
                if (self._graph_cur, self._graph_parity) != (self._cur, self._parity):
                    self._capture()          # buffer and parity out of phase with the recording: record again
            while n >= 2:
                self._graph.replay()
                n -= 2
        for _ in range(n):
            self._advance()
        return self

    def _capture(self):
        """Record two consecutive steps (A->B, B->A) in one CUDA graph."""
        torch.cuda.synchronize()
        keep = [b.clone() for b in self._bufs]
        ib_keep = None
        if self.ib is not None:
            ib_keep = [t.clone() for t in (self._marker_u, self._marker_force)] + (
                [self._body_dev.clone()] if self._body_dev is not None else [])
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):   # warm-up outside capture
            self._advance(); self._advance()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        self._graph_cur = self._cur
        self._graph_parity = self._parity
        with torch.cuda.graph(g):
            self._advance(); self._advance()
        torch.cuda.synchronize()
        for b, k in zip(self._bufs, keep):   # capture does not execute, but the warm-up did: restore
            b.copy_(k)
        if ib_keep is not None:
            self._marker_u.copy_(ib_keep[0]); self._marker_force.copy_(ib_keep[1])
            if self._body_dev is not None:
                self._body_dev.copy_(ib_keep[2])
        self._graph = g


class Ensemble:
    """Independent simulations advanced together on one GPU, each with its rigid-body ODE on the host
    (vsb_run_host_ode_multi) -- e.g. the reduced-velocity sweep of a VIV study, every case being one run of the
    reference's examples/2d/vortex_induced_vibration.py.  One host thread serves all bodies: whichever domain's IB
    force has arrived gets its Newmark update (dyn.py:5-51) and its next step, so the device always has other
    domains' kernels to run while one waits for the host.

        ens = Ensemble([Stepper(spec_k, body=body_k, dyn_mode="host").set_f(f0_k) for k in range(8)])
        ens.step(1000)
        d, h = ens.steppers[3].body_state()
    """

    def __init__(self, steppers):
        self.steppers = list(steppers)
        if not 1 <= len(self.steppers) <= 64:
            raise ValueError("an Ensemble holds 1..64 steppers")
        dev = self.steppers[0].device
        for st in self.steppers:
            if not isinstance(st, Stepper) or st.device != dev:
                raise ValueError("all members must be Steppers on the same device")
            if not st._host_ode_eligible():
                raise ValueError("every member needs an immersed body with dyn_mode='host', overlap on, no halo and "
                                 "face operations the fused wall blocks support")
        self._mains = [torch.cuda.Stream(device=dev) for _ in self.steppers]

    def step(self, n=1):
        """Advance every member n reference time steps."""
        n = int(n)
        if n <= 0:
            return self
        kinds = {st._kind for st in self.steppers}
        if None in kinds:
            raise L.VsbError("no state loaded: call set_f(f) on every member first")
        if len(kinds) != 1:
            raise L.VsbError("members are in different states: step them to the same point first")
        if kinds == {"F"}:              # prologue S_0 = collide(F_0), counted as the first step like Stepper.step
            for st in self.steppers:
                st.step(1)
            n -= 1
            if n == 0:
                return self
        cur = torch.cuda.current_stream()
        k = len(self.steppers)
        for st, main in zip(self.steppers, self._mains):
            main.wait_stream(cur)
            st._host_ode_prepare(main)
        args = (C.POINTER(L.VsbStepArgs) * k)(*[C.pointer(st._args) for st in self.steppers])
        mdfs = (C.POINTER(L.VsbMdfArgs) * k)(*[C.pointer(st._mdf) for st in self.steppers])
        params = (C.POINTER(L.VsbBodyParams) * k)(*[C.pointer(st._hparams) for st in self.steppers])
        pinned = (C.c_void_p * k)(*[st._body_pin.data_ptr() for st in self.steppers])
        plans = (C.POINTER(L.VsbHostPlan) * k)(*[C.pointer(st._plan) for st in self.steppers])
        rc = L.lib().vsb_run_host_ode_multi(k, args, mdfs, params, pinned, plans, n)
        for st, main in zip(self.steppers, self._mains):
            cur.wait_stream(main)
        L.check(rc)
        for st in self.steppers:
            st._host_ode_finish(n)
            st.n_steps += n
        return self
