"""Benchmark of the fused IB-LBM time step (BASELINE.json metric: MLUPS and % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload: BASELINE config 2 (C2) -- D2Q9 BGK, 1024 x 1024, a 512-marker elastically mounted cylinder, MDF with 5
iterations + Guo forcing, inlet NEBB / outlet equilibrium.  MLUPS = lattice cells x time steps / seconds / 1e6.

One bench STEP is one *chunk*: CHUNK = 500 lattice time steps of every domain of the ensemble -- the reference's own
unit of work, one `update_chunk(carry, CHUNK_STEPS)` call between two host synchronisations
(examples/2d/vortex_induced_vibration.py:28,150-157,200-203).  A single lattice time step of a 1024^2 domain lasts
~14 us, far too short to be a timed unit; whatever --steps is, the timed loop consists of whole CUDA-graph replays
(`graph_replays` / `eager_steps` in the JSON line say what ran).

* value      device-resident throughput.  The 75 MB working set of one domain fits the 126 MB L2, so the timed loop
             rotates over 8 independent domains (an ensemble, e.g. a reduced-velocity sweep) whose combined working
             set exceeds 4x L2: every step streams its populations from and to HBM.  The single-domain, L2-resident
             figure is reported next to it (config.single_domain).
* e2e        the same chunks through the public API from HOST buffers: pinned f -> device, K chunks with the rigid-body
             ODE on the host (per time step the device posts the body force into a 16-byte host mailbox and the
             92-byte body state comes back, as north_star prescribes), per chunk the (d, h) record is read on the host
             like the reference's driver loop does, f -> host at the end.
* roofline   the fused kernel (vsb_step) alone, CUDA-event timed, 72 B/cell algorithmic traffic against the
             measured HBM copy bandwidth of MEASURED_PEAKS.json; `traffic` = DRAM read + write bytes of one launch
             from the committed ncu capture (profiles/r02_traffic.json), null when there is none.
* also       BASELINE configs 1, 3, 4 and 5 on one GPU.
* cpu_baseline / --impl reference   the C + OpenMP restatement of the reference's composed step
             (oracle/c, kind "port": jax is not installable so the reference's JAX-CPU backend cannot run)
             on the box's host cores, thread count set explicitly to the cores this process may use.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MLUPS (IB-LBM step, D2Q9 1024x1024 VIV cylinder, 512 markers, MDF + Guo)"
WORKLOAD = "C2: D2Q9 BGK IB-LBM VIV cylinder 1024x1024, 512 markers, MDF(5) + Guo forcing"
L2_BYTES = 126e6
CHUNK = 500          # lattice time steps per domain per bench step (reference CHUNK_STEPS, vortex_induced_vibration.py:28)
GRAPH_STEPS = 10     # lattice time steps per domain per CUDA-graph replay (even: the buffers ping-pong)
# ensemble members: overlap the IB chain with the bulk inside each domain (the domains overlap each other either way)
ENSEMBLE_OVERLAP = os.environ.get("VSB_BENCH_OVERLAP", "1") != "0"
# how the MDF iterations of an ensemble member are chained (profiles/r02_chain_modes_*.txt, 8 domains, two visits): all
# iterations in one cooperative launch with grid barriers 77.2 / 78.6 GLUPS, one launch per iteration 75.7 / 78.1, the
# marker-space cluster kernel 73.5 / 76.5; enqueuing the chain before the bulk (chain_first) is within the noise
ENSEMBLE_CHAIN = os.environ.get("VSB_BENCH_CHAIN", "barrier")
CHAIN_FIRST = os.environ.get("VSB_BENCH_CHAIN_FIRST", "0") != "0"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """DRAM bytes of one launch of the dominant kernel from the committed ncu capture (or None)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as fh:
            t = json.load(fh)
        return float(t["dram_read_bytes"]) + float(t["dram_write_bytes"]), t
    except Exception:
        return None, None


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE, text=True, bufsize=1)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, t0, t1):
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 8]
        if not rows:
            return None
        sm = sorted(float(r[1]) for r in rows)
        reasons = [n for i, n in ((4, "hw_slowdown"), (5, "hw_thermal_slowdown"), (6, "sw_thermal_slowdown"),
                                  (7, "sw_power_cap")) if any(r[i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": reasons, "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows)}


# ------------------------------------------------------------------------------ CPU side (oracle; baseline only)
def cpu_workload():
    """BASELINE config 2 for the C port: same geometry / parameters as vivsim_b200.configs.viv_cylinder_2d()."""
    import numpy as np
    from oracle import recipes
    from vivsim_b200 import configs
    spec, body = configs.viv_cylinder_2d()
    f0 = recipes.uniform_init(spec)
    return spec, body, np.ascontiguousarray(f0)


class CpuArm:
    """The reference's unfused step restated in C + OpenMP (oracle/c/iblbm_ref.c) on the host cores.  The OpenMP thread
    count is set explicitly to the cores this process may run on: torchrun exports OMP_NUM_THREADS=1, which must not
    shrink the baseline."""

    def __init__(self):
        from oracle import cport
        self.spec, body, f0 = cpu_workload()
        self.runner = cport.CRunner(self.spec, f0, body=body)
        self.threads = host_cores()
        self.cells = self.spec["shape"][0] * self.spec["shape"][1]
        self.runner.run(1, threads=self.threads)

    def time_steps(self, n):
        t = time.perf_counter()
        self.runner.run(n, threads=self.threads)
        return time.perf_counter() - t

    def describe(self, n, what):
        return (f"{what}: {n} lattice time steps of one 1024x1024 VIV-cylinder domain (C2), C + OpenMP restatement of the "
                f"reference's unfused step (oracle/c/iblbm_ref.c), {self.threads} threads")


def time_cpu(budget_s):
    arm = CpuArm()
    per = arm.time_steps(2) / 2
    n = max(3, int(budget_s / max(per, 1e-6)))
    dt = arm.time_steps(n)
    return {"value": arm.cells * n / dt / 1e6, "unit": "MLUPS", "cores": arm.threads, "kind": "port",
            "sample": arm.describe(n, "bounded sample")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = CpuArm()
    per = arm.time_steps(2) / 2
    # one reference step = a bounded sample of the chunk our arm's step consists of (CHUNK time steps of 8 domains would
    # take ~40 s per step here): n time steps of one domain, ~2 s of CPU work per step
    n = max(2, min(CHUNK, int(2.0 / max(per, 1e-6))))
    for _ in range(args.warmup):
        arm.time_steps(n)
    dt = sum(arm.time_steps(n) for _ in range(args.steps))
    value = arm.cells * n * args.steps / dt / 1e6
    base = {"value": value, "unit": "MLUPS", "cores": arm.threads, "kind": "port",
            "sample": arm.describe(n, f"each of the {args.steps} timed steps")}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "MLUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "lattice_steps_per_bench_step": n,
                       "note": "reference = vivsim's algorithm restated in C + OpenMP on the host cores (jax/jaxlib are "
                               "not installable in this image, so the reference's own JAX CPU backend cannot run); a step "
                               "is a bounded sample of the chunk the GPU arm's step consists of; MLUPS does not depend "
                               "on the sample length"},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------ GPU side
def timed(fn, sync):
    import torch
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record()
    fn()
    e1.record()
    sync()
    return e0.elapsed_time(e1) * 1e-3, w0, time.time()


def build_graph(steppers, steps_each):
    """One CUDA graph that advances every stepper `steps_each` steps.  The domains are independent simulations, so each
    gets its own stream inside the graph: one domain's immersed-boundary chain overlaps another domain's bulk pass."""
    import torch
    streams = [torch.cuda.Stream() for _ in steppers]

    def enqueue():
        main = torch.cuda.current_stream()
        for s, st in zip(steppers, streams):
            st.wait_stream(main)
            with torch.cuda.stream(st):
                s.advance_raw(steps_each)
        for st in streams:
            main.wait_stream(st)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        enqueue()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        enqueue()
    torch.cuda.synchronize()
    g._streams = streams
    return g


class GraphLoop:
    """Timed loop made of whole graph replays only: `run(n)` advances every member n x steps_each time steps."""

    def __init__(self, steppers, steps_each):
        self.steppers, self.steps_each = steppers, steps_each
        self.graph = build_graph(steppers, steps_each)
        self.replays = 0

    def run(self, n_replays):
        for _ in range(int(n_replays)):
            self.graph.replay()
        self.replays += int(n_replays)


def cells_of(spec):
    n = 1
    for k in spec["shape"]:
        n *= k
    return n


def extra_workload(name, hbm_gbs):
    """Single-GPU runs of the other BASELINE configurations, reported next to the headline."""
    import torch
    from vivsim_b200 import Stepper, configs
    sync = torch.cuda.synchronize
    if name == "c1":
        # BASELINE config 1: README lid-driven cavity, 100 x 100, BGK, NEE on four walls, 1000 steps.  The four walls
        # interact at the corners, so the step takes the ordered wall fix-up (several tiny launches): launch-bound.
        spec, _ = configs.cavity()
        st = Stepper(spec, use_graph=True)
        import numpy as np
        from vivsim_b200 import lbm
        shape = tuple(spec["shape"])
        f0 = lbm.get_equilibrium(torch.ones(shape, device="cuda"), torch.zeros((2,) + shape, device="cuda"))
        st.set_f(f0)
        st.step(21)
        n = 1000
        dt, _, _ = timed(lambda: st.step(n), sync)
        u = st.macroscopic()[1]
        return {"workload": "C1: D2Q9 BGK lid-driven cavity 100x100, NEE on four walls, 1000 steps (README example)",
                "mlups": cells_of(spec) * n / dt / 1e6, "steps": n, "ms_per_step": dt / n * 1e3,
                "launches_per_step": st.n_launch_per_step,
                "note": "40 KB per population set: resident in L2, bound by launch latency (CUDA graph of 2 steps)",
                "finite": bool(torch.isfinite(u).all()), "lid_velocity_seen": float(u[0, :, -1].mean())}
    if name == "c3":
        spec, body = configs.sphere_3d()
        label, bpc, steps = "C3: D3Q19 KBC IB-LBM sphere 256^3, 2562 markers, MDF(3) + EDM", 152, 40
    elif name == "c4":
        spec, body = configs.viv_cylinder_2d_large()
        label, bpc, steps = "C4: D2Q9 KBC VIV cylinder 16384^2, 3276 markers, MDF(5) + EDM (single GPU)", 72, 12
    else:
        spec, body = configs.oscillating_cylinder_3d()
        label, bpc, steps = ("C5: D3Q19 MRT oscillating cylinder 1024x512x512, 695570 markers, MDF(3) + Guo-MRT "
                             "(single GPU)"), 152, 12
    st = Stepper(spec, body=body, dyn_mode="device", follow=2 if name == "c5" else 1) if body else Stepper(spec)
    st.set_f(configs.uniform_state(spec, noise=1e-3))
    st.step(3)
    loop = GraphLoop([st], 2)
    loop.run(1)
    dt, _, _ = timed(lambda: loop.run(steps // 2), sync)
    mlups = cells_of(spec) * steps / dt / 1e6
    ok = bool(torch.isfinite(st.state).all())
    del st, loop
    torch.cuda.empty_cache()
    return {"workload": label, "mlups": mlups, "steps": steps, "ms_per_step": dt / steps * 1e3,
            "hbm_frac_of_measured": mlups * 1e6 * bpc / (hbm_gbs * 1e9), "finite": ok}


def multi_gpu_config(name, world, rank, hbm_gbs):
    """BASELINE configs 4 and 5 on all N GPUs (strong scaling of one fixed global grid, collective: every rank calls it).
    C4: 16384^2 KBC VIV cylinder, slabs along x, the body's chain on the rank that owns it.  C5: 1024 x 512 x 512 MRT
    with the 695 k-marker cylinder, slabs along x, the IB chain SHARED by all ranks over peer memory (ib='shard')."""
    import torch
    import torch.distributed as dist
    from vivsim_b200 import configs
    from vivsim_b200.multidevice import SlabStepper
    if name == "c4":
        spec, body = configs.viv_cylinder_2d_large()
        label, bpc, steps, kw = "C4: D2Q9 KBC VIV cylinder 16384^2, 3276 markers, MDF(5) + EDM", 72, 24, dict(ib="owner", follow=1)
    else:
        spec, body = configs.oscillating_cylinder_3d()
        label, bpc, steps, kw = ("C5: D3Q19 MRT oscillating cylinder 1024x512x512, 695570 markers, MDF(3) + Guo-MRT",
                                 152, 24, dict(ib="shard", follow=2))
    st = SlabStepper(spec, body=dict(body), dyn_mode="device", halo="peer", **kw)
    st.set_f_local(configs.uniform_state(dict(spec, shape=st.slab.local_shape), noise=1e-3))
    st.step(3)
    loop = GraphLoop([st], 2)
    loop.run(2)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    dt, _, _ = timed(lambda: loop.run(steps // 2), lambda: (torch.cuda.synchronize(), dist.barrier(), torch.cuda.synchronize()))
    t = torch.tensor([dt], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t)
    ok = torch.tensor([1.0 if bool(torch.isfinite(st.stepper.state).all()) else 0.0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    timed_out = bool(st.peer is not None and st.peer.timed_out()) or bool(st.ib_shard is not None and st.ib_shard.timed_out())
    mlups = cells_of(spec) * steps / dt / 1e6
    out = {"workload": f"{label} on {world} GPUs (strong scaling, x-slabs, peer-memory halo)", "ib": st.ib_mode,
           "mlups": mlups, "steps": steps, "ms_per_step": dt / steps * 1e3,
           "per_gpu_hbm_frac_of_measured": mlups / world * 1e6 * bpc / (hbm_gbs * 1e9), "finite": bool(ok.item()),
           "timed_out": timed_out, "launches_per_step": st.n_launch_per_step}
    del loop, st
    torch.cuda.empty_cache()
    return out


def parity_vs_one_gpu(world, rank):
    """N slabs against one GPU on the same global problem (C2 recipe with walls and an immersed cylinder, KBC periodic
    case): max relative difference of F after 20 steps, gathered on every rank, compared on rank 0."""
    import torch
    import torch.distributed as dist
    from vivsim_b200 import Stepper, configs
    from vivsim_b200.multidevice import SlabStepper
    out = {}
    nx, ny = 128 * world, 128
    spec, _ = configs.viv_cylinder_2d(nx=nx, ny=ny, n_marker=64, radius=8.0, u0=0.08, nu=0.02, moving=False,
                                      center=(nx - 64.0, ny / 2))
    f0 = configs.uniform_state(spec, noise=1e-3)
    dist.broadcast(f0, 0)
    s = SlabStepper(spec).set_f_global(f0)
    s.step(20)
    got = s.gather_f()
    if rank == 0:
        ref = Stepper(spec).set_f(f0)
        ref.step(20)
        r = ref.get_f()
        out = {"case": f"C2 recipe {nx}x{ny}, 64-marker cylinder, 20 steps, halo={s.halo}",
               "max_rel_diff": float((got - r).abs().max() / r.abs().max()), "bit_exact": bool(torch.equal(got, r))}
    del s
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from vivsim_b200 import Stepper, configs, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.lib()
    hbm_gbs, peak_src = peaks()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    spec, body = configs.viv_cylinder_2d()
    cells = cells_of(spec)
    bytes_per_domain = 2 * 9 * 4 * cells
    n_rep = int(-(-4 * L2_BYTES // bytes_per_domain)) + 1          # ensemble working set > 4x L2
    f0 = configs.uniform_state(spec, noise=1e-3)
    steppers = []
    K, W = args.steps, args.warmup
    replays_per_step = CHUNK // GRAPH_STEPS
    if world == 1:
        for i in range(n_rep):
            st = Stepper(spec, body=dict(body), dyn_mode="device", overlap=ENSEMBLE_OVERLAP, ib_chain=ENSEMBLE_CHAIN,
                         chain_first=CHAIN_FIRST)
            st.set_f(f0)
            st.step(1)     # prologue: internal state is now S_0
            steppers.append(st)
        multi = None
    else:
        # weak scaling: every ensemble member is one (world x 1024) x 1024 channel cut into 1024-wide slabs, one
        # elastically mounted cylinder per slab; after every step the populations crossing the cuts are exchanged
        from vivsim_b200.multidevice import SlabStepper
        nxl = spec["shape"][0]
        gspec = dict(spec, shape=(nxl * world, spec["shape"][1]))
        gspec.pop("ib")

        def local_ib(slab):
            sp, _ = configs.viv_cylinder_2d(center=(slab.x0 + nxl / 2, spec["shape"][1] / 2))
            return sp["ib"]

        f_loc = torch.cat([f0[:, -1:], f0, f0[:, :1]], dim=1).contiguous()     # local slab + periodic ghost layers
        for i in range(n_rep):
            st = SlabStepper(gspec, local_ib=local_ib, body=dict(body), dyn_mode="device", halo="peer", ib_chain=ENSEMBLE_CHAIN)
            st.set_f_local(f_loc)
            st.step(1)
            steppers.append(st)
        multi = {"decomposition": f"{world} slabs of {nxl} x {spec['shape'][1]} along x per ensemble member",
                 "exchange": ("peer-mapped symmetric memory over NVLink, inside the step's CUDA graph, no NCCL on the data "
                              "path: interior rows start at once; on a second stream ONE launch of the fused kernel waits "
                              "for the neighbours' flag words, updates the two edge rows, stores the 3 crossing "
                              "populations (4 KB each per side) into the neighbours' ghost rows from its epilogue and "
                              "publishes the step (VsbStepArgs.halo)" if steppers[0].stepper.halo_fused else
                              "peer-mapped symmetric memory over NVLink, inside the step's CUDA graph, no NCCL on the data "
                              "path: interior rows start at once; a second stream waits for the neighbours' flag words "
                              "(vsb_halo_wait), updates the two edge rows and stores the 3 crossing populations' edge rows "
                              "(4 KB each) into the neighbours' ghost rows (vsb_halo_send)"),
                 "halo_fused": bool(steppers[0].stepper.halo_fused),
                 "halo_bytes_per_step_per_rank": steppers[0].slab.halo_bytes_per_step()}
    loop = GraphLoop(steppers, GRAPH_STEPS)
    loop.run(W * replays_per_step)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        time.sleep(0.3)
    loop.replays = 0
    dt, w0, w1 = timed(lambda: loop.run(K * replays_per_step), sync)
    timed_replays = loop.replays
    if world > 1:
        t = torch.tensor([dt], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t)
    lattice_steps = K * CHUNK                      # per domain
    value = cells * lattice_steps * n_rep * world / dt / 1e6

    clocks = sampler.summary(w0, w1) if sampler else None
    clock_note = None
    if sampler and (clocks is None or clocks["samples"] < 3) and world == 1:
        # timed region shorter than a few 100 ms sampling periods (tiny --steps): sample an identical follow-up loop
        t_a = time.time()
        while time.time() - t_a < 1.0:
            loop.run(replays_per_step)
            torch.cuda.synchronize()
        clocks = sampler.summary(t_a, time.time())
        clock_note = "timed region shorter than the sampling period; sampled during an identical untimed follow-up loop"
    if sampler:
        sampler.stop()

    inner = [s.stepper if world > 1 else s for s in steppers]
    finite = all(bool(torch.isfinite(s.state).all()) for s in inner)
    if world > 1 and any(s.peer is not None and s.peer.timed_out() for s in steppers):
        raise RuntimeError("a halo wait timed out: the ranks did not run the same number of steps")
    launches_per_lattice_step = inner[0].n_launch_per_step

    # ---- e2e through the public API from HOST buffers
    f_host = f0.cpu().pin_memory()
    state_bytes = f_host.numel() * 4
    if world == 1:
        # host-side rigid-body ODE as north_star prescribes.  The workload is the one `value` is measured on: n_rep
        # independent domains through the public Ensemble API; one host thread serves every body's ODE
        # (vsb_run_host_ode_multi); after every chunk the host reads the (d, h) record of the chunk, as the
        # reference's driver loop does (vortex_induced_vibration.py:200-205).
        from vivsim_b200 import Ensemble
        del loop
        hist_body = dict(body, history=CHUNK)
        ens = Ensemble([Stepper(spec, body=dict(hist_body), dyn_mode="host", ib_chain=ENSEMBLE_CHAIN) for _ in range(n_rep)])
        for st in ens.steppers:
            st.set_f(f_host)
        ens.step(40)
        for st in ens.steppers:
            st.get_f()
        backs = [torch.empty_like(f_host).pin_memory() for _ in ens.steppers]      # page-locked result buffers
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for st in ens.steppers:
            st.set_f(f_host)
        records = []
        for _ in range(K):
            ens.step(CHUNK)
            records.append([st.body_history(CHUNK - 1) for st in ens.steppers])
        for b, st in zip(backs, ens.steppers):
            b.copy_(st.get_f(), non_blocking=True)
        torch.cuda.synchronize()
        te = time.perf_counter() - t0
        assert all(bool(torch.isfinite(b).all()) for b in backs)
        import numpy as np
        assert all(np.isfinite(d).all() and np.isfinite(h).all() for rec in records for d, h in rec)
        e2e_steps = K * CHUNK
        del ens, backs
        ke = 2 * CHUNK
        # for context: ONE domain alone (every step waits for its own host round trip, nothing else to run meanwhile)
        st = Stepper(spec, body=dict(body), dyn_mode="host")
        st.set_f(f_host); st.step(5); st.get_f()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st.set_f(f_host)
        st.step(ke)
        f_back1 = st.get_f().to("cpu", non_blocking=False)
        torch.cuda.synchronize()
        te1 = time.perf_counter() - t0
        assert bool(torch.isfinite(f_back1).all())
        e2e_single = {"value": cells * ke / te1 / 1e6, "unit": "MLUPS", "lattice_steps": ke,
                      "note": "one domain alone through Stepper.step (vsb_run_host_ode): every time step waits for its "
                              "own device -> host -> device round trip; populations host -> device and back included"}
        del st
        # for context: the same with the ODE on the device (what the reference does inside its jitted scan):
        # host transfers only at the chunk boundaries
        st = Stepper(spec, body=dict(body), dyn_mode="device", use_graph=True)
        st.set_f(f_host); st.step(6); st.get_f()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st.set_f(f_host)
        st.step(ke)
        f_back2 = st.get_f().to("cpu", non_blocking=False)
        st.body_state()
        torch.cuda.synchronize()
        te2 = time.perf_counter() - t0
        assert bool(torch.isfinite(f_back2).all())
        e2e_device_ode = {"value": cells * ke / te2 / 1e6, "unit": "MLUPS", "lattice_steps": ke,
                          "note": "one domain, rigid-body ODE on the device: f host->device, graph-replayed steps, "
                                  "f and body state device->host"}
        del st
        e2e = {"value": cells * e2e_steps * n_rep / te / 1e6, "unit": "MLUPS", "steps": K,
               "lattice_steps_per_domain": e2e_steps,
               # per bench step (= chunk of CHUNK time steps of n_rep domains)
               "h2d_bytes_per_step": n_rep * (state_bytes / K + CHUNK * _lib.BODY_BYTES),
               "d2h_bytes_per_step": n_rep * (state_bytes / K + CHUNK * 16),
               "regime": "compute-bound: populations cross PCIe once per domain and direction in the whole run "
                         f"({n_rep} x {state_bytes / 1e6:.1f} MB each way), the per-time-step traffic is the body "
                         "mailbox / state",
               "note": (f"the {n_rep}-domain ensemble `value` is measured on, through the public API (Ensemble.step -> "
                        "vsb_run_host_ode_multi) from pinned HOST buffers: per domain f -> device once, then per time "
                        "step the device posts the body force into a 16-byte host mailbox, one host thread polls all "
                        "mailboxes, advances that body's Newmark ODE on the CPU, sends the 92-byte body state back and "
                        "enqueues its next step; per chunk the host reads every domain's (d, h) record; f -> host once "
                        "per domain at the end -- all inside the timed region"),
               "single_domain_host_ode": e2e_single, "single_domain_device_ode": e2e_device_ode}
    else:
        # N GPUs: pinned slab -> device, K chunks with the peer-memory halo exchange (body ODE on the device, one
        # cylinder per slab), body state -> host per chunk, slab -> host at the end
        f_loc_host = torch.cat([f_host[:, -1:], f_host, f_host[:, :1]], dim=1).contiguous().pin_memory()
        sync()
        t0 = time.perf_counter()
        for st in steppers:
            st.set_f_local(f_loc_host)
            st.step(2)      # prologue + one step: the state is back in the buffer (and IB parity) the graph starts from
        for _ in range(K):
            loop.run(replays_per_step)
            for st in steppers:
                st.stepper.body_state()
        f_back = [st.get_f_local().to("cpu", non_blocking=False) for st in steppers]
        sync()
        te = time.perf_counter() - t0
        t = torch.tensor([te], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        te = float(t)
        assert all(bool(torch.isfinite(f).all()) for f in f_back)
        loc_bytes = f_loc_host.numel() * 4
        e2e = {"value": cells * (K * CHUNK + 2) * n_rep * world / te / 1e6, "unit": "MLUPS", "steps": K,
               "lattice_steps_per_domain": K * CHUNK + 2,
               "h2d_bytes_per_step": world * n_rep * loc_bytes / K,
               "d2h_bytes_per_step": world * n_rep * (loc_bytes / K + _lib.BODY_BYTES),
               "note": f"the {n_rep}-member ensemble `value` is measured on (every member one slab-decomposed channel): "
                       "per rank and member pinned slab -> device once, K chunks of graph-replayed steps with the "
                       "peer-memory halo exchange and the body ODE on the device, every body's state -> host per chunk, "
                       "slab -> host once per member; max over ranks"}
        del f_back, loop

    parity = None
    also_multi = []
    if world > 1:
        parity = parity_vs_one_gpu(world, rank)
        if not args.no_extra:
            del steppers, inner
            torch.cuda.empty_cache()
            for name in ("c4", "c5"):
                try:
                    also_multi.append(multi_gpu_config(name, world, rank, hbm_gbs))
                except Exception as exc:  # report, never hide
                    also_multi.append({"workload": name, "error": f"{type(exc).__name__}: {exc}"})

    line = None
    if rank == 0:
        single = None
        if world == 1:
            # ---- ONE domain (what BASELINE config 2 names literally): 75 MB working set, L2-resident
            one = GraphLoop(steppers[:1], GRAPH_STEPS)
            one.run(20)
            n1 = 4 * replays_per_step
            dt1, _, _ = timed(lambda: one.run(n1), torch.cuda.synchronize)
            single = {"mlups": cells * n1 * GRAPH_STEPS / dt1 / 1e6, "us_per_lattice_step": dt1 / (n1 * GRAPH_STEPS) * 1e6,
                      "note": "one 1024^2 domain alone: both population buffers (75 MB) stay in the 126 MB L2, so the "
                              "step is bound by the latency of the IB chain, not by HBM"}
            del one

        # ---- roofline of the dominant kernel: vsb_step alone (no IB, no wall kernels), rotating buffers; same kernel
        # instantiation and runtime flags as in the step (Guo forcing enabled, force zero outside the IB window,
        # which covers 1 % of the cells)
        roofline = None
        if world == 1:
            plain = dict(spec); plain.pop("ib"); plain["post"] = []
            ks = [Stepper(plain).set_f(f0) for _ in range(n_rep)]
            for s in ks:
                s.step(1)
            gk = GraphLoop(ks, 2)
            gk.run(3)
            nrk = 40
            dtk, _, _ = timed(lambda: gk.run(nrk), torch.cuda.synchronize)
            nk = nrk * 2 * n_rep
            per_launch = dtk / nk
            achieved = 72.0 * cells / per_launch / 1e9
            traffic, traffic_src = ncu_traffic()
            roofline = {"bound": "hbm", "kernel": "vsb::k_step<2, BGK, vec4> (fused pull-stream + moments + BGK + Guo)",
                        "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs,
                        "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                        "us_per_launch": per_launch * 1e6, "algorithmic_bytes_per_launch": 72 * cells,
                        "note": "72 B/cell (9 x 4 B read + 9 x 4 B write) x 1048576 cells per launch; CUDA events over "
                                f"{nk} launches rotating over {n_rep} domains (working set {n_rep * bytes_per_domain / 1e6:.0f} MB > L2)"}
            del ks, gk

        also = list(also_multi)
        if not args.no_extra and world == 1:
            del steppers, inner
            torch.cuda.empty_cache()
            for name in ("c1", "c3", "c4", "c5"):
                try:
                    also.append(extra_workload(name, hbm_gbs))
                except Exception as exc:  # report, never hide
                    also.append({"workload": name, "error": f"{type(exc).__name__}: {exc}"})

        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cpu = time_cpu(budget_s=12.0)

        line = {"metric": METRIC, "value": value, "unit": "MLUPS", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD + ", 2-DOF Newmark body on device"
                                       + (f"; per GPU, {world} slabs per channel" if world > 1 else ""),
                           "bench_step": f"one chunk = {CHUNK} lattice time steps of each of the {n_rep} ensemble domains"
                                         + (" on every GPU" if world > 1 else "")
                                         + " (reference update_chunk, vortex_induced_vibration.py:28,150-157)",
                           "lattice_steps_per_bench_step": CHUNK, "cells_per_lattice_step": cells,
                           "ensemble_domains": n_rep, "ib_chain": ENSEMBLE_CHAIN, "chain_first": CHAIN_FIRST,
                           "graph_replays": timed_replays, "eager_steps": 0,
                           "lattice_steps_per_graph_replay": GRAPH_STEPS,
                           "us_per_lattice_step": dt / (lattice_steps * n_rep) * 1e6,
                           "l2": f"{n_rep} independent domains rotated so the working set ({n_rep * bytes_per_domain / 1e6:.0f} MB) "
                                 "exceeds 4x L2: inputs come from HBM every step (no L2 flush needed)",
                           "single_domain": single,
                           "hbm_frac_of_measured": value / world * 1e6 * 72 / (hbm_gbs * 1e9),
                           "multi_gpu": multi, "parity_vs_1gpu": parity,
                           "state_finite": finite},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": launches_per_lattice_step * lattice_steps * n_rep,
                "clocks": dict(clocks or {}, **({"note": clock_note} if clock_note else {})),
                "also": also}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the C1 / C3 / C4 / C5 single-GPU measurements")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    args.steps = max(args.steps, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
