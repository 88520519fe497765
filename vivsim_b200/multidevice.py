"""Slab domain decomposition across GPUs with halo exchange over NCCL / NVLink.

Replaces the reference's ``vivsim/multidevice.py`` (``stream_cross_devices``: four ``lax.ppermute`` of
the three populations that cross each cut, 2-D only, no caller).  Design from the physics:

* the global grid is cut along x, the slowest axis, so the layer ``f[q, x_edge]`` of every population is
  one contiguous block (NY floats in 2-D, NY*NZ in 3-D): no pack / unpack kernels are needed;
* every rank stores its slab with one ghost layer on each side, local shape ``(nx_local + 2, NY[, NZ])``,
  physical rows ``[1, nx_local + 1)``;
* after each fused step only the populations moving across a cut are exchanged with the two ring
  neighbours (periodic wrap): ``c_x = +1`` populations go right, ``c_x = -1`` go left -- 3 of 9 in D2Q9,
  5 of 19 in D3Q19 -- as one batch of NCCL send/recv pairs (``ncclGroupStart/End``);
* wall boundary operations on the x faces are applied by the rank that owns that face; y / z faces and the
  obstacle mask are applied by every rank on its rows; an immersed body is owned by the rank whose slab
  contains its window, and the total IB force is combined with a small all-reduce.

The exchange is written on torch tensors, so the same code runs on CPU tensors over gloo (tests) and on
CUDA tensors over NCCL (production).  There is no data-path collective besides the neighbour exchange.
"""

import os

import numpy as np
import torch
import torch.distributed as dist

RIGHT_MOVING = {2: (1, 5, 8), 3: (1, 7, 9, 11, 13)}   # c_x = +1  (lbm/lattice.py:58, lbm3d/lattice.py:43-58)
LEFT_MOVING = {2: (3, 7, 6), 3: (2, 8, 10, 12, 14)}   # c_x = -1


class Slab:
    """Geometry of one rank's slab of a global grid cut along x."""

    def __init__(self, global_shape, rank, world):
        self.global_shape = tuple(int(n) for n in global_shape)
        self.dim = len(self.global_shape)
        self.rank, self.world = int(rank), int(world)
        nx = self.global_shape[0]
        if nx % self.world:
            raise ValueError(f"NX = {nx} must be divisible by the number of ranks ({self.world})")
        self.nx_local = nx // self.world
        if self.nx_local < 4:
            raise ValueError("each slab needs at least 4 x-layers")
        self.x0 = self.rank * self.nx_local                       # global x of local row 1
        self.local_shape = (self.nx_local + 2,) + self.global_shape[1:]
        self.rows = (1, self.nx_local + 1)
        self.left = (self.rank - 1) % self.world
        self.right = (self.rank + 1) % self.world
        self.owns_left_wall = self.rank == 0
        self.owns_right_wall = self.rank == self.world - 1

    def to_local_x(self, x_global):
        return x_global - self.x0 + 1

    def scatter(self, field_global):
        """Local slab (with ghost layers filled periodically) of a global array whose axis 1 is x."""
        nx = self.global_shape[0]
        idx = (np.arange(self.x0 - 1, self.x0 + self.nx_local + 1)) % nx
        if isinstance(field_global, torch.Tensor):
            return field_global.index_select(1, torch.as_tensor(idx, device=field_global.device)).contiguous()
        return np.ascontiguousarray(np.take(field_global, idx, axis=1))

    def halo_bytes_per_step(self):
        face = int(np.prod(self.global_shape[1:]))
        return 2 * len(RIGHT_MOVING[self.dim]) * face * 4


def exchange_halo(state, slab, group=None):
    """Fill the ghost layers of ``state`` (Q, nx_local + 2, ...) from the ring neighbours.

    Row ``nx_local`` (last physical) of the right-moving populations goes to the right neighbour's ghost
    row 0; row 1 of the left-moving populations goes to the left neighbour's ghost row ``nx_local + 1``.
    Returns the list of outstanding requests (already waited on for world == 1)."""
    n = slab.nx_local
    if slab.world == 1:
        for q in RIGHT_MOVING[slab.dim]:
            state[q, 0].copy_(state[q, n])
        for q in LEFT_MOVING[slab.dim]:
            state[q, n + 1].copy_(state[q, 1])
        return []
    ops = []
    for q in RIGHT_MOVING[slab.dim]:
        ops.append(dist.P2POp(dist.isend, state[q, n], slab.right, group))
        ops.append(dist.P2POp(dist.irecv, state[q, 0], slab.left, group))
    for q in LEFT_MOVING[slab.dim]:
        ops.append(dist.P2POp(dist.isend, state[q, 1], slab.left, group))
        ops.append(dist.P2POp(dist.irecv, state[q, n + 1], slab.right, group))
    return dist.batch_isend_irecv(ops)


def wait_all(reqs):
    for r in reqs:
        r.wait()


class PeerHalo:
    """Ghost-layer exchange through peer-mapped symmetric memory (vsb_halo_push): the state buffers of all ranks are
    allocated with torch.distributed._symmetric_memory, so a kernel on this GPU stores straight into the neighbours'
    ghost layers over NVLink and signals with a flag word; no NCCL call, graph-capturable."""

    def __init__(self, slab, q, group=None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        from . import _lib as L
        group = group if group is not None else dist.group.WORLD
        dev = torch.device("cuda", torch.cuda.current_device())
        self.slab = slab
        self.buf = symm.empty((2, q) + slab.local_shape, dtype=torch.float32, device=dev)
        self.buf.zero_()
        self.flags = symm.empty((16,), dtype=torch.int32, device=dev)
        self.flags.zero_()
        self._h_buf = symm.rendezvous(self.buf, group)
        self._h_flags = symm.rendezvous(self.flags, group)
        torch.cuda.synchronize()
        dist.barrier(group)                       # every rank's flags are zero before anyone can signal
        self.counter = torch.zeros(4, dtype=torch.int32, device=dev)
        nbytes = self.buf[0].numel() * 4
        bp, fp = self._h_buf.buffer_ptrs, self._h_flags.buffer_ptrs
        me, left, right = slab.rank, slab.left, slab.right
        self.args = []
        for i in range(2):
            a = L.VsbHaloArgs()
            a.grid = L.grid_of(slab.local_shape)
            a.state = bp[me] + i * nbytes
            a.left_state = bp[left] + i * nbytes
            a.right_state = bp[right] + i * nbytes
            a.my_flags, a.left_flags, a.right_flags = fp[me], fp[left], fp[right]
            a.counter = self.counter.data_ptr()
            self.args.append(a)
        self._C, self._L = C, L

    def buffers(self):
        return [self.buf[0], self.buf[1]]

    def push(self, index, st=None):
        """Send the edge layers of buffer `index` and wait for the neighbours' (stream `st` or the current one)."""
        L, C = self._L, self._C
        L.check(L.lib().vsb_halo_push(C.byref(self.args[index]), st if st is not None else L.stream()))

    def send(self, index, st=None):
        """Copy the edge layers of buffer `index` into the neighbours' ghost layers and publish the step number."""
        L, C = self._L, self._C
        L.check(L.lib().vsb_halo_send(C.byref(self.args[index]), st if st is not None else L.stream()))

    def wait(self, st=None):
        """Wait until both neighbours have published the current step number."""
        L, C = self._L, self._C
        L.check(L.lib().vsb_halo_wait(C.byref(self.args[0]), st if st is not None else L.stream()))

    def timed_out(self):
        return bool(self.counter[2].item())

    def check(self):
        """Synchronise and raise VsbError if a wait for a neighbour gave up (vsb_sync_status)."""
        L, C = self._L, self._C
        L.check(L.lib().vsb_sync_status(C.c_void_p(self.counter.data_ptr()), L.stream()))


class LocalHalo:
    """Single-rank stand-in for PeerHalo: the "neighbours" are the slab itself (periodic wrap), so `send` copies the
    edge layers into the slab's own ghost layers and `wait` has nothing to wait for.  Exercises the pipelined step
    (interior rows / wait / edge rows / send) on one GPU."""

    def __init__(self, slab, buffers):
        self.slab, self._bufs = slab, buffers

    def buffers(self):
        return self._bufs

    def _copy(self, index, st):
        stream = torch.cuda.ExternalStream(st.value) if st is not None and st.value else torch.cuda.current_stream()
        with torch.cuda.stream(stream):
            exchange_halo(self._bufs[index], self.slab)

    def push(self, index, st=None):
        self._copy(index, st)

    def send(self, index, st=None):
        self._copy(index, st)

    def wait(self, st=None):
        pass

    def timed_out(self):
        return False

    def check(self):
        pass


class LocalFusedHalo(LocalHalo):
    """LocalHalo whose hand-shake runs inside the edge-row launch (VsbStepArgs.halo): the slab is its own left and
    right neighbour, so the launch's peer stores land in its own ghost layers and its flag words are its own.  Puts
    the fused wait / send / publish code of k_step under test on one GPU."""

    def __init__(self, slab, buffers):
        import ctypes as C
        from . import _lib as L
        super().__init__(slab, buffers)
        dev = buffers[0].device
        self.flags = torch.zeros(16, dtype=torch.int32, device=dev)
        self.counter = torch.zeros(4, dtype=torch.int32, device=dev)
        self.args = []
        for buf in buffers:
            a = L.VsbHaloArgs()
            a.grid = L.grid_of(slab.local_shape)
            a.state = a.left_state = a.right_state = buf.data_ptr()
            a.my_flags = a.left_flags = a.right_flags = self.flags.data_ptr()
            a.counter = self.counter.data_ptr()
            self.args.append(a)
        self._C, self._L = C, L

    def timed_out(self):
        return bool(self.counter[2].item())

    def check(self):
        L, C = self._L, self._C
        L.check(L.lib().vsb_sync_status(C.c_void_p(self.counter.data_ptr()), L.stream()))


# Cost model behind the marker shares (measured on B200 with BASELINE config 5, profiles/r02_summary.md): one marker costs a
# rank about 0.4 ns per MDF iteration; a rank whose slab holds reachable window cells also computes their velocity and
# collides them after the chain, about 0.2 ns per cell.
COST_PER_MARKER_ITER = 0.4e-9
COST_PER_WINDOW_CELL = 0.2e-9


def balance_marker_shares(n_markers, n_iter, cells_per_rank):
    """Fractions of the markers per rank that equalise (window work of the rank) + (marker work of the rank): ranks
    whose slabs hold the body's window get fewer markers, down to none (water filling).  Host logic."""
    extra = np.asarray(cells_per_rank, dtype=np.float64) * COST_PER_WINDOW_CELL
    total = float(n_markers) * n_iter * COST_PER_MARKER_ITER
    world = extra.size
    if total <= 0:
        return np.full(world, 1.0 / world)
    order = np.argsort(extra)
    level = 0.0
    for k in range(world, 0, -1):            # the k ranks with the least window work share the markers
        level = (total + extra[order[:k]].sum()) / k
        if level >= extra[order[k - 1]]:
            break
    share = np.maximum(level - extra, 0.0)
    return share / share.sum()


def plan_ib_shards(markers, window, world, dense, moving_in_window=False, margin=2, shares=None):
    """Divide the markers of one immersed body among `world` ranks (host logic, NumPy only).

    The ranks share the multi-direct-forcing chain (csrc/vsb_ibshard.cu): rank r handles a contiguous range of the
    stored marker list and needs, in its copy of the window fields, every cell its markers' 4-point stencils can touch
    -- its NEED BOX.  Markers are split into equal-count groups along the axis of the body's largest extent (for the
    z-aligned cylinder of BASELINE config 5 that is z: every rank gets a slice of the cylinder), so the boxes are compact
    and neighbouring boxes overlap by a few cells only.

    dense: the tiled kernel is used -- within each group the markers are ordered and cut into chunks by
    stepper.cut_marker_chunks.  moving_in_window: the body moves relative to the window (follow = 0, or it rotates),
    so every rank needs the whole window; otherwise the window follows the body and `margin` cells cover the
    sub-cell drift.  shares: fraction of the markers per rank (default: equal).
    Returns dict(perm, marker_ranges (world, 2), chunk_offsets | None, chunk_ranges (world, 2), need_lo, need_hi
    (world, dim) window-local [lo, hi), axis, cells: ascending flat window indices of the cells a stencil can reach)."""
    from .stepper import cut_marker_chunks
    markers = np.asarray(markers, dtype=np.float32)
    n, dim = markers.shape
    origin, size = window
    origin = np.floor(np.asarray(origin, dtype=np.float64)).astype(np.int64)
    size = np.asarray(size, dtype=np.int64)
    axis = int(np.argmax(markers.max(axis=0) - markers.min(axis=0))) if n else 0
    order = np.argsort(markers[:, axis], kind="stable")
    if shares is None:
        bounds = [(n * r) // world for r in range(world + 1)]
    else:                                     # uneven shares (balance_marker_shares)
        cum = np.concatenate([[0.0], np.cumsum(np.asarray(shares, dtype=np.float64))])
        bounds = [int(round(n * c / cum[-1])) for c in cum]
        bounds[0], bounds[-1] = 0, n
    perm_parts, marker_ranges, chunk_ranges, offsets = [], [], [], [0]
    pos = 0
    for r in range(world):
        grp = np.sort(order[bounds[r]:bounds[r + 1]])          # the caller's order within a group
        if dense and grp.size:
            sub_perm, sub_off = cut_marker_chunks(markers[grp])
            grp = grp[sub_perm]
            c0 = len(offsets) - 1
            offsets.extend((pos + sub_off[1:]).tolist())
            chunk_ranges.append((c0, len(offsets) - 1))
        else:
            chunk_ranges.append((len(offsets) - 1, len(offsets) - 1))
        perm_parts.append(grp)
        marker_ranges.append((pos, pos + grp.size))
        pos += grp.size
    perm = np.concatenate(perm_parts) if perm_parts else np.zeros(0, dtype=np.int64)
    need_lo = np.zeros((world, dim), dtype=np.int64)
    need_hi = np.zeros((world, dim), dtype=np.int64)
    for r, (b, e) in enumerate(marker_ranges):
        if e == b:
            continue                                          # empty share: empty box
        if moving_in_window:
            need_hi[r] = size
            continue
        base = np.floor(markers[perm[b:e]].astype(np.float64) - origin).astype(np.int64)
        need_lo[r] = np.clip(base.min(axis=0) - 1 - margin, 0, size)
        need_hi[r] = np.clip(base.max(axis=0) + 3 + margin, 0, size)
    # window cells some marker's stencil can reach: base - 1 .. base + 2, and the base itself drifts by at most one cell
    # either way while the window follows the body
    if moving_in_window or n == 0:
        cells = np.arange(int(np.prod(size)), dtype=np.int32)
    else:
        from .stepper import reachable_window_cells
        cells = reachable_window_cells(markers, origin, size)
    return dict(perm=perm, marker_ranges=np.asarray(marker_ranges, dtype=np.int64), cells=cells,
                chunk_offsets=np.asarray(offsets, dtype=np.int32) if dense else None,
                chunk_ranges=np.asarray(chunk_ranges, dtype=np.int64), need_lo=need_lo, need_hi=need_hi, axis=axis)


class IbShard:
    """Everything a rank needs to take part in the shared IB chain: the marker plan and the peer-mapped window fields,
    flag words and sum slots (torch.distributed._symmetric_memory), packed into a VsbIbShard."""

    def __init__(self, slab, ib, n_iter, moving_in_window, group=None):
        import torch.distributed._symmetric_memory as symm
        from . import _lib as L
        group = group if group is not None else dist.group.WORLD
        dev = torch.device("cuda", torch.cuda.current_device())
        dim = slab.dim
        if slab.world > L.MAX_RANKS:
            raise ValueError(f"the shared IB chain supports at most {L.MAX_RANKS} ranks")
        origin, size = ib["window"]
        self.win_size = tuple(int(n) for n in size)
        markers = np.asarray(ib["markers"], dtype=np.float32)
        wcells = int(np.prod(self.win_size))
        dense = dim == 3 and markers.shape[0] * 4 ** dim > 2 * wcells and markers.shape[0] > 480
        # first pass: which window cells are reachable, and how many of them lie in each slab -> marker shares
        probe = plan_ib_shards(markers, ib["window"], 1, False, moving_in_window)
        cx = probe["cells"].astype(np.int64) // int(np.prod(self.win_size[1:])) + int(np.floor(origin[0]))
        per_rank = np.bincount(np.clip(cx // slab.nx_local, 0, slab.world - 1), minlength=slab.world)
        self.shares = balance_marker_shares(markers.shape[0], n_iter, per_rank)
        self.plan = plan_ib_shards(markers, ib["window"], slab.world, dense, moving_in_window, shares=self.shares)
        nc = 2 if dim == 2 else 4
        self.fields = symm.empty((2, n_iter + 1) + self.win_size + (nc,), dtype=torch.float32, device=dev)
        self.fields.zero_()
        self.flags = symm.empty((16,), dtype=torch.int32, device=dev)
        self.flags.zero_()
        self.sums = symm.empty((L.MAX_RANKS * 4,), dtype=torch.float32, device=dev)
        self.sums.zero_()
        # staging slots [source rank] for the force field, and the local field the fluid kernels read
        self.staging = symm.empty((slab.world,) + self.win_size + (nc,), dtype=torch.float32, device=dev)
        self.staging.zero_()
        self.force_field = torch.zeros(self.win_size + (nc,), dtype=torch.float32, device=dev)
        self.cells = torch.as_tensor(self.plan["cells"], device=dev)
        handles = [symm.rendezvous(t, group) for t in (self.fields, self.flags, self.sums, self.staging)]
        torch.cuda.synchronize()
        dist.barrier(group)
        self.counter = torch.zeros(4, dtype=torch.int32, device=dev)
        sh = L.VsbIbShard()
        sh.n_ranks, sh.rank = slab.world, slab.rank
        for r in range(slab.world):
            sh.fields[r], sh.flags[r], sh.sums[r], sh.staging[r] = (h.buffer_ptrs[r] for h in handles)
            for d in range(dim):
                sh.need_lo[r][d] = int(self.plan["need_lo"][r][d])
                sh.need_hi[r][d] = int(self.plan["need_hi"][r][d])
            sh.x_lo[r] = r * slab.nx_local
            sh.x_hi[r] = (r + 1) * slab.nx_local
        sh.marker_begin, sh.marker_end = (int(x) for x in self.plan["marker_ranges"][slab.rank])
        sh.chunk_begin, sh.chunk_end = (int(x) for x in self.plan["chunk_ranges"][slab.rank])
        sh.counter = self.counter.data_ptr()
        sh.force_field = self.force_field.data_ptr()
        sh.cells, sh.n_cells = self.cells.data_ptr(), int(self.cells.numel())
        self.ev_window_done = torch.cuda.Event()
        self.ev_window_done.record()              # materialise the cudaEvent_t handle
        sh.ev_window_done = self.ev_window_done.cuda_event
        self.trace = None
        if os.environ.get("VSB_SHARD_TRACE"):       # timing aid: %globaltimer at entry / exit of every flag barrier
            self.trace = torch.zeros(8192, dtype=torch.int64, device=dev)
            sh.trace = self.trace.data_ptr()
        self.args = sh
        self.slab = slab
        self._handles = handles

    def timed_out(self):
        return bool(self.counter[2].item())

    def check(self):
        """Synchronise and raise VsbError if a flag barrier of the shared chain gave up (vsb_sync_status)."""
        import ctypes as C
        from . import _lib as L
        L.check(L.lib().vsb_sync_status(C.c_void_p(self.counter.data_ptr()), L.stream()))


def localize_spec(spec, slab, local_ib=None, ib_mode="owner"):
    """Per-rank step description: local extent with ghost layers, x-face operations only on the owning rank,
    immersed body (``spec['ib']`` or ``local_ib(slab)``, global coordinates) shifted to local coordinates."""
    out = dict(spec)
    out["shape"] = slab.local_shape
    post = []
    for item in spec.get("post", ()):
        if item[0] == "mask":
            m = item[1]
            m = m.detach().cpu().numpy() if isinstance(m, torch.Tensor) else np.asarray(m)
            post.append(("mask", slab.scatter(m[None])[0]))
            continue
        loc = item[1]
        if loc == "left" and not slab.owns_left_wall:
            continue
        if loc == "right" and not slab.owns_right_wall:
            continue
        kw = dict(item[2]) if len(item) > 2 else {}
        for k, v in kw.items():
            if hasattr(v, "ndim") and v.ndim > 0 and loc not in ("left", "right"):
                arr = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
                kw[k] = slab.scatter(arr[None])[0]          # face arrays of y / z faces run along x
        post.append((item[0], loc, kw))
    out["post"] = post
    ib = local_ib(slab) if local_ib is not None else spec.get("ib")
    out["ib"] = None
    if ib is not None and ib_mode == "shard":
        out["ib"] = dict(ib)                  # global coordinates: every rank takes part (IbShard)
    elif ib is not None:
        (ox, *orest), size = ib["window"]
        lo, hi = int(np.floor(ox)), int(np.floor(ox)) + int(size[0])
        inside = lo >= slab.x0 + 2 and hi <= slab.x0 + slab.nx_local - 2
        overlaps = hi > slab.x0 and lo < slab.x0 + slab.nx_local
        if overlaps and not inside:
            raise ValueError(f"the IB window x-range [{lo}, {hi}) must lie at least 2 layers inside one slab "
                             f"(rank {slab.rank} owns [{slab.x0}, {slab.x0 + slab.nx_local})) for ib='owner'; "
                             "use ib='shard' (or 'auto'), which shares the chain among all ranks")
        if inside:
            loc_ib = dict(ib)
            markers = np.array(ib["markers"], dtype=np.float32, copy=True)
            markers[:, 0] = markers[:, 0] - np.float32(slab.x0) + np.float32(1)
            loc_ib["markers"] = markers
            loc_ib["window"] = ((slab.to_local_x(ox),) + tuple(orest), tuple(size))
            out["ib"] = loc_ib
    g = spec.get("g")
    if isinstance(g, torch.Tensor) and g.ndim == slab.dim + 1:
        out["g"] = slab.scatter(g)
    return out


class SlabStepper:
    """One rank's part of a slab-decomposed simulation: a ``Stepper`` on the local extent plus the halo exchange after
    every step.  Collective: every rank must call ``step`` / ``advance_raw`` with the same count.

    halo = "peer": ghost layers are written directly into the neighbours' buffers by vsb_halo_push (symmetric memory,
    NVLink; graph-capturable).  halo = "nccl": torch.distributed send/recv of the edge layers.  "auto": peer when the
    symmetric-memory rendezvous succeeds, else NCCL."""

    def __init__(self, spec, rank=None, world=None, group=None, local_ib=None, body=None, halo="auto", ib="auto",
                 halo_fused=None, **kw):
        """halo_fused: None (default: on unless VSB_HALO_FUSED=0) / False -- whether the edge-row launch of the pipelined
        pass does the neighbour hand-shake itself (one launch) or vsb_halo_wait / vsb_halo_send run around it (three).
        ib: how an immersed body of the global spec is distributed -- "owner": the rank whose slab contains the IB
        window (>= 2 layers from a cut) computes the whole chain; "shard": the markers are divided among all ranks and
        the window fields are shared through peer memory (a body may sit on or move across a cut; the chain of a
        large body is spread over all GPUs); "auto": "owner" when the window of a fixed body fits one slab, else
        "shard".  Bodies given per slab through local_ib are always "owner"."""
        from .stepper import Stepper
        self.group = group
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.slab = Slab(spec["shape"], rank, world)
        if ib not in ("auto", "owner", "shard"):
            raise ValueError("ib must be 'auto', 'owner' or 'shard'")
        gib = spec.get("ib") if local_ib is None else None
        if gib is None or world == 1:
            ib = "owner"
        elif ib == "auto":
            (ox, *_), size = gib["window"]
            lo, hi = int(np.floor(ox)), int(np.floor(ox)) + int(size[0])
            k = lo // self.slab.nx_local
            fits = lo >= k * self.slab.nx_local + 2 and hi <= (k + 1) * self.slab.nx_local - 2
            ib = "owner" if (fits and body is None) else "shard"
        self.ib_mode = ib
        self.local_spec = localize_spec(spec, self.slab, local_ib, ib)
        has_body = self.local_spec["ib"] is not None
        self.ib_shard = None
        if ib == "shard":
            if not torch.cuda.is_available():
                raise RuntimeError("ib='shard' shares the IB chain through peer-mapped GPU memory")
            follow = int(kw.get("follow", 1))
            moving_in_window = body is not None and (follow == 0 or bool(body.get("rotation", False)))
            self.ib_shard = IbShard(self.slab, gib, int(gib.get("n_iter", 5)), moving_in_window, group)
            kw = dict(kw, ib_shard=self.ib_shard, dyn_mode=kw.get("dyn_mode", "device"))
        self.peer = None
        self.halo_error = None
        if world > 1 and halo in ("auto", "peer"):
            try:
                self.peer = PeerHalo(self.slab, 9 if self.slab.dim == 2 else 19, group)
            except Exception as exc:   # noqa: BLE001 -- symmetric memory unavailable: fall back to NCCL unless forced
                if halo == "peer":
                    raise
                self.halo_error = f"{type(exc).__name__}: {exc}"
        self.halo = "peer" if self.peer is not None else ("nccl" if world > 1 else "local")
        buffers = self.peer.buffers() if self.peer is not None else None
        self.stepper = Stepper(self.local_spec, rows=self.slab.rows, body=body if has_body else None, buffers=buffers, **kw)
        self.owns_body = has_body
        if world == 1 and halo in ("pipelined-local", "fused-local"):   # one rank, but through the same pipelined pass as peer mode
            self.peer = (LocalHalo if halo == "pipelined-local" else LocalFusedHalo)(self.slab, self.stepper._bufs)
            self.halo = halo
        if self.peer is not None:
            self.stepper.attach_halo(self.peer, fused=halo_fused)   # the halo kernels become part of every pass of the stepper
        self.n_launch_per_step = self.stepper.n_launch_per_step

    # -- state in / out (reference convention F)
    def set_f_global(self, f_global):
        self.stepper.set_f(self.slab.scatter(f_global))
        return self

    def set_f_local(self, f_local_with_ghosts):
        self.stepper.set_f(f_local_with_ghosts)
        return self

    def get_f_local(self):
        """Physical rows of F_n on this rank, shape (Q, nx_local, ...)."""
        if self.peer is not None and self.stepper._kind == "S":
            self.peer.wait()                      # the neighbours' last sends into this rank's ghost layers
        self.check()
        return self.stepper.get_f()[:, 1:-1].contiguous()

    def check(self):
        """Raise VsbError if any cross-GPU wait of the steps taken so far timed out (synchronises this rank's stream).
        Called by every method that hands results to the host."""
        if self.peer is not None:
            self.peer.check()
        if self.ib_shard is not None:
            self.ib_shard.check()

    def gather_f(self):
        """Global F_n on every rank (all-gather along x) -- for tests and I/O."""
        loc = self.get_f_local()
        if self.slab.world == 1:
            return loc
        parts = [torch.empty_like(loc) for _ in range(self.slab.world)]
        dist.all_gather(parts, loc, group=self.group)
        return torch.cat(parts, dim=1)

    def total_force(self):
        """Sum over all bodies / ranks of the hydrodynamic force on the bodies (small all-reduce)."""
        st = self.stepper
        dim = st.dim
        self.check()
        # sharded chain: every rank holds the forces of its share of the markers (the rest are zero)
        h = (-st.marker_force.sum(dim=0)) if self.owns_body else torch.zeros(dim, device=st.device)
        if self.slab.world > 1:
            dist.all_reduce(h, group=self.group)
        return h

    def marker_force(self):
        """+F on every marker in the caller's order, on every rank (all-reduce of the ranks' shares when the chain is
        shared; zeros on ranks that do not own the body otherwise)."""
        st = self.stepper
        if not self.owns_body:
            f = None
        else:
            f = st.marker_force.clone()
        if self.slab.world > 1:
            n = torch.tensor([0 if f is None else f.shape[0]], device=st.device)
            dist.all_reduce(n, op=dist.ReduceOp.MAX, group=self.group)
            if f is None:
                f = torch.zeros((int(n), st.dim), device=st.device)
            dist.all_reduce(f, group=self.group)
        return f

    # -- stepping
    def _exchange(self):
        if self.peer is None:                     # peer mode: the stepper's passes already contain the halo kernels
            wait_all(exchange_halo(self.stepper.state, self.slab, self.group))

    def advance_raw(self, n=1):
        """n x (fused step + halo exchange) on the current stream; graph-capturable with halo == "peer"."""
        for _ in range(int(n)):
            self.stepper.advance_raw(1)
            self._exchange()
        return self

    def step(self, n=1):
        st = self.stepper
        st._require_state()
        for _ in range(int(n)):
            if st._kind == "F":
                # the F state came with periodic ghost layers (scatter) or the caller filled them
                st.step(1)          # prologue S_0 = collide(F_0) on the physical rows
            else:
                st.advance_raw(1)
            self._exchange()
        return self
