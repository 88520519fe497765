"""Benchmark of the fused IB-LBM time step (BASELINE.json metric: MLUPS and % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one IB-LBM time step of one 1024 x 1024 D2Q9 domain with a 512-marker elastically mounted
cylinder (BASELINE config 1: MDF 5 iterations + Guo forcing, BGK, inlet NEBB / outlet equilibrium).
MLUPS = cells x steps / seconds / 1e6.

* value      device-resident throughput.  The 75 MB working set of one domain fits the 126 MB L2, so the
             timed loop rotates over several independent domains (an ensemble) whose combined working set
             exceeds 4x L2: every step streams its populations from and to HBM.  The single-domain,
             L2-resident figure is reported next to it (config.l2_resident_mlups).
* e2e        the same workload through the public API from HOST buffers: pinned f -> device, K steps with
             the rigid-body ODE on the host (one 72-byte device->host read of the body force and one
             72-byte host->device write of the kinematics per step, as north_star prescribes), f -> host.
* roofline   the fused kernel (vsb_step) alone, CUDA-event timed, 72 B/cell algorithmic traffic against the
             measured HBM copy bandwidth of MEASURED_PEAKS.json.
* cpu_baseline / --impl reference   the C + OpenMP restatement of the reference's composed step
             (oracle/c, kind "port": jax is not installable so the reference's JAX-CPU backend cannot run)
             on the box's host cores.
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
L2_BYTES = 126e6
# ensemble members: overlap the IB chain with the bulk inside each domain (the domains overlap each other either way)
ENSEMBLE_OVERLAP = os.environ.get("VSB_BENCH_OVERLAP", "1") != "0"
# dram__bytes_read.sum + dram__bytes_write.sum of one k_step<2,BGK,vec4> launch at 1024^2 (ncu --set full,
# profiles/r01_ncu_full_k_step_c2_final.txt: the 1181-block launch of the bench command)
TRAFFIC_NCU = 39.0e6   # 37.71 MB read + 1.28 MB written to DRAM during the launch (the rest of the writes leave L2 later)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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
    """BASELINE config 1 for the C port: same geometry / parameters as vivsim_b200.configs.viv_cylinder_2d()."""
    import numpy as np
    from oracle import recipes
    from vivsim_b200 import configs
    spec, body = configs.viv_cylinder_2d()
    f0 = recipes.uniform_init(spec)
    return spec, body, np.ascontiguousarray(f0)


def time_cpu(budget_s, max_steps=None):
    from oracle import cport
    spec, body, f0 = cpu_workload()
    r = cport.CRunner(spec, f0, body=body)
    threads = cport.num_threads()
    r.run(1)
    t = time.perf_counter(); r.run(2); per = (time.perf_counter() - t) / 2
    n = max(3, int(budget_s / max(per, 1e-6)))
    if max_steps:
        n = min(n, max_steps)
    t = time.perf_counter(); r.run(n); dt = time.perf_counter() - t
    cells = spec["shape"][0] * spec["shape"][1]
    return {"value": cells * n / dt / 1e6, "unit": "MLUPS", "cores": threads, "kind": "port",
            "sample": f"{n} time steps of the 1024x1024 VIV-cylinder workload, C + OpenMP restatement of the reference's "
                      f"unfused step (oracle/c/iblbm_ref.c), {threads} threads"}, dt / n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, per = time_cpu(budget_s=90.0, max_steps=max(args.steps, 3))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "MLUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2: D2Q9 BGK IB-LBM VIV cylinder 1024x1024, 512 markers, MDF(5) + Guo forcing",
                       "note": "reference = vivsim's algorithm restated in C + OpenMP on the host cores (jax/jaxlib are "
                               "not installable in this image, so the reference's own JAX CPU backend cannot run)"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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


def run_loop(graph, steppers, steps_each, n_steps):
    """Advance exactly n_steps (summed over the ensemble)."""
    per_replay = len(steppers) * steps_each
    for _ in range(n_steps // per_replay):
        graph.replay()
    rem = n_steps % per_replay
    i = 0
    while rem > 0:
        k = min(steps_each, rem)
        steppers[i % len(steppers)].advance_raw(k)
        rem -= k
        i += 1


def extra_workload(name, steps, hbm_gbs):
    """Single-GPU HBM-bound configurations reported next to the headline (BASELINE configs 2 and 3)."""
    import torch
    from vivsim_b200 import Stepper, configs
    if name == "c3":
        spec, body = configs.sphere_3d()
        label, bpc = "C3: D3Q19 KBC IB-LBM sphere 256^3, 2562 markers, MDF(3) + EDM", 152
    else:
        spec, body = configs.viv_cylinder_2d_large()
        label, bpc = "C4: D2Q9 KBC VIV cylinder 16384^2, 3276 markers, MDF(5) + EDM (single GPU)", 72
    st = Stepper(spec, body=body, dyn_mode="device") if body else Stepper(spec)
    st.set_f(configs.uniform_state(spec, noise=1e-3))
    st.step(3)
    g = build_graph([st], 2)
    sync = torch.cuda.synchronize
    dt, _, _ = timed(lambda: run_loop(g, [st], 2, steps), sync)
    cells = 1
    for n in spec["shape"]:
        cells *= n
    mlups = cells * steps / dt / 1e6
    ok = bool(torch.isfinite(st.state).all())
    del st, g
    torch.cuda.empty_cache()
    return {"workload": label, "mlups": mlups, "steps": steps, "ms_per_step": dt / steps * 1e3,
            "hbm_frac_of_measured": mlups * 1e6 * bpc / (hbm_gbs * 1e9), "finite": ok}


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
    cells = spec["shape"][0] * spec["shape"][1]
    bytes_per_domain = 2 * 9 * 4 * cells
    n_rep = int(-(-4 * L2_BYTES // bytes_per_domain)) + 1          # ensemble working set > 4x L2
    f0 = configs.uniform_state(spec, noise=1e-3)
    steppers = []
    steps_each = 8      # steps per domain per graph replay (measured: 2 -> 15.7, 4 -> 14.7, 8 -> 14.35 us/step)
    K, W = args.steps, args.warmup
    if world == 1:
        for i in range(n_rep):
            st = Stepper(spec, body=dict(body), dyn_mode="device", overlap=ENSEMBLE_OVERLAP)
            st.set_f(f0)
            st.step(1)     # prologue: internal state is now S_0
            steppers.append(st)
        graph = build_graph(steppers, steps_each)
        loop = lambda n: run_loop(graph, steppers, steps_each, n)
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
            st = SlabStepper(gspec, local_ib=local_ib, body=dict(body), dyn_mode="device")
            st.set_f_local(f_loc)
            st.step(1)
            steppers.append(st)

        peer = all(s.halo == "peer" for s in steppers)
        if peer:    # halo kernels are part of the step graph: same replay loop as on one GPU
            graph = build_graph(steppers, steps_each)
            loop = lambda n: run_loop(graph, steppers, steps_each, n)
            exchange = ("peer-mapped symmetric memory over NVLink, inside the step's CUDA graph, no NCCL on the data path: "
                        "interior rows start at once; a second stream waits for the neighbours' flag words "
                        "(vsb_halo_wait), updates the two edge rows and stores the 3 crossing populations' edge rows "
                        "(4 KB each) into the neighbours' ghost rows (vsb_halo_send)")
        else:
            graph = None

            def loop(n):
                for k in range(n):
                    steppers[k % n_rep].step(1)

            exchange = ("NCCL send/recv of the 3 populations crossing each cut after every step (eager; symmetric "
                        f"memory unavailable: {steppers[0].halo_error})")
        multi = {"decomposition": f"{world} slabs of {nxl} x {spec['shape'][1]} along x per ensemble member",
                 "exchange": exchange, "halo_bytes_per_step_per_rank": steppers[0].slab.halo_bytes_per_step()}
    loop(W)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        time.sleep(0.3)
    dt, w0, w1 = timed(lambda: loop(K), sync)
    if world > 1:
        t = torch.tensor([dt], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t)
    value = cells * K * world / dt / 1e6

    clocks = sampler.summary(w0, w1) if sampler else None
    clock_note = None
    if world > 1 and dt < 0.5:
        # short timed region on N GPUs: every rank runs the same untimed follow-up loop (~1.5 s; dt is the max over ranks,
        # so the step count is identical everywhere) while rank 0 samples the clocks
        per_replay = len(steppers) * steps_each
        n_follow = int(min(1.5 / max(dt / K, 1e-7), 2e6))
        n_follow = max(n_follow - n_follow % per_replay, per_replay)        # whole graph replays only
        t_a = time.time()
        loop(n_follow)
        sync()
        if sampler:
            clocks = sampler.summary(t_a, time.time())
            clock_note = "timed region shorter than the sampling period; sampled during an identical untimed follow-up loop"
    elif sampler and (clocks is None or clocks["samples"] < 3):
        # the timed region is shorter than the 100 ms sampling period: sample an identical follow-up loop
        t_a = time.time()
        while time.time() - t_a < 1.5 and world == 1:
            loop(2 * len(steppers) * 50)
            torch.cuda.synchronize()
        clocks = sampler.summary(t_a, time.time())
        clock_note = "timed region shorter than the sampling period; sampled during an identical untimed follow-up loop"
    if sampler:
        sampler.stop()

    # ---- e2e through the public API from HOST buffers
    f_host = f0.cpu().pin_memory()
    state_bytes = f_host.numel() * 4
    ke = max(10, min(K, 3000))
    if world == 1:
        # host-side rigid-body ODE as north_star prescribes: every step one device->host read of the body force and
        # one host->device write of the kinematics (16-byte mailbox up, 92-byte body state down, every step)
        # The workload is the one `value` is measured on: n_rep independent domains (an ensemble, e.g. a reduced-velocity
        # sweep), through the public Ensemble API: one host thread serves every body's ODE (vsb_run_host_ode_multi).
        from vivsim_b200 import Ensemble
        ens = Ensemble([Stepper(spec, body=dict(body), dyn_mode="host") for _ in range(n_rep)])
        for st in ens.steppers:
            st.set_f(f_host)
        ens.step(5)
        for st in ens.steppers:
            st.get_f()
        backs = [torch.empty_like(f_host).pin_memory() for _ in ens.steppers]      # page-locked result buffers
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for st in ens.steppers:
            st.set_f(f_host)
        ens.step(ke)
        for b, st in zip(backs, ens.steppers):
            b.copy_(st.get_f(), non_blocking=True)
        torch.cuda.synchronize()
        te = time.perf_counter() - t0
        f_back = backs[0]
        assert all(bool(torch.isfinite(b).all()) for b in backs)
        e2e_domains = n_rep
        del ens, backs
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
        e2e_single = {"value": cells * ke / te1 / 1e6, "unit": "MLUPS", "steps": ke,
                      "note": "one domain alone through Stepper.step (vsb_run_host_ode): every step waits for its own "
                              "device -> host -> device round trip"}
        per_step_io = _lib.BODY_BYTES
        per_step_up = 16                      # VsbHostMail: force[3] + seq written by the device into pinned host memory
        e2e_note = (f"the {n_rep}-domain ensemble `value` is measured on, through the public API (Ensemble.step -> "
                    "vsb_run_host_ode_multi) from pinned HOST buffers: per domain f -> device once, then per step the "
                    "device posts the body force into a 16-byte host mailbox, one host thread polls all mailboxes, "
                    "advances that body's Newmark ODE on the CPU, sends the 92-byte body state back and enqueues its "
                    "next step (every step, inside the timed region); f -> host once per domain")
        del st
        # for context: the same chunk with the ODE on the device (what the reference does inside its jitted scan):
        # host transfers only at the chunk boundaries
        st = Stepper(spec, body=dict(body), dyn_mode="device", use_graph=True)
        st.set_f(f_host); st.step(6); st.get_f()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st.set_f(f_host)
        st.step(ke)
        f_back2 = st.get_f().to("cpu", non_blocking=False)
        body_back = st.body_state()
        torch.cuda.synchronize()
        te2 = time.perf_counter() - t0
        e2e_device_ode = {"value": cells * ke / te2 / 1e6, "unit": "MLUPS", "steps": ke,
                          "note": "same chunk, rigid-body ODE on the device: f host->device, ke graph-replayed steps, "
                                  "f and body state device->host"}
        del st
    else:
        # N GPUs: pinned slab -> device, ke steps with halo exchange (body ODE on the device), slab -> host
        f_loc_host = torch.cat([f_host[:, -1:], f_host, f_host[:, :1]], dim=1).contiguous().pin_memory()
        st = steppers[0]
        sync()
        t0 = time.perf_counter()
        st.set_f_local(f_loc_host)
        st.step(ke)
        f_back = st.get_f_local().to("cpu", non_blocking=False)
        sync()
        te = time.perf_counter() - t0
        t = torch.tensor([te], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        te = float(t)
        per_step_io = 0
        per_step_up = 0
        e2e_note = ("per rank: pinned slab -> device once, steps with NCCL halo exchange and the body ODE on the device, "
                    "slab -> host once; max over ranks")
    assert bool(torch.isfinite(f_back).all())
    e2e = {"value": cells * ke * world * (e2e_domains if world == 1 else 1) / te / 1e6, "unit": "MLUPS",
           "steps": ke * (e2e_domains if world == 1 else 1),
           "h2d_bytes_per_step": state_bytes / ke + per_step_io, "d2h_bytes_per_step": state_bytes / ke + per_step_up,
           "note": e2e_note}
    if world == 1:
        e2e["single_domain_host_ode"] = e2e_single
        e2e["chunked_device_ode"] = e2e_device_ode

    line = None
    if rank == 0:
        l2_mlups = None
        if world == 1:
            # ---- single domain, L2-resident
            g1 = build_graph(steppers[:1], steps_each)
            n1 = max(200, min(K, 20000))
            run_loop(g1, steppers[:1], steps_each, 50)
            dt1, _, _ = timed(lambda: run_loop(g1, steppers[:1], steps_each, n1), torch.cuda.synchronize)
            l2_mlups = cells * n1 / dt1 / 1e6

        # ---- roofline of the dominant kernel: vsb_step alone (no IB, no wall kernels), rotating buffers; same kernel
        # instantiation and runtime flags as in the step (Guo forcing enabled, force zero outside the IB window,
        # which covers 1 % of the cells)
        plain = dict(spec); plain.pop("ib"); plain["post"] = []
        ks = [Stepper(plain).set_f(f0) for _ in range(n_rep)]
        for s in ks:
            s.step(1)
        gk = build_graph(ks, 2)
        nk = 2 * n_rep * 20
        run_loop(gk, ks, 2, 2 * n_rep * 3)
        dtk, _, _ = timed(lambda: run_loop(gk, ks, 2, nk), torch.cuda.synchronize)
        per_launch = dtk / nk
        achieved = 72.0 * cells / per_launch / 1e9
        roofline = {"bound": "hbm", "kernel": "vsb::k_step<2, BGK, vec4> (fused pull-stream + moments + BGK + Guo)",
                    "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs,
                    "traffic": TRAFFIC_NCU, "peak_source": peak_src, "us_per_launch": per_launch * 1e6,
                    "algorithmic_bytes_per_launch": 72 * cells,
                    "note": "72 B/cell (9 x 4 B read + 9 x 4 B write) x 1048576 cells per launch; CUDA events over "
                            f"{nk} launches rotating over {n_rep} domains (working set {n_rep * bytes_per_domain / 1e6:.0f} MB > L2); "
                            "traffic = dram read + write bytes of one launch from the committed ncu --set full capture "
                            "(profiles/): the 37.7 MB of writes mostly stay in the 126 MB L2 until later launches evict them"}
        del ks, gk

        also = []
        if not args.no_extra and world == 1:
            for name, n in (("c3", 40), ("c4", 12)):
                try:
                    also.append(extra_workload(name, n, hbm_gbs))
                except Exception as exc:  # report, never hide
                    also.append({"workload": name, "error": f"{type(exc).__name__}: {exc}"})

        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cpu, _ = time_cpu(budget_s=12.0)

        inner = [s.stepper if world > 1 else s for s in steppers]
        finite = all(bool(torch.isfinite(s.state).all()) for s in inner)
        if world > 1 and any(s.peer is not None and s.peer.timed_out() for s in steppers):
            raise RuntimeError("a halo wait timed out: the ranks did not run the same number of steps")
        line = {"metric": METRIC, "value": value, "unit": "MLUPS", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "C2: D2Q9 BGK IB-LBM VIV cylinder 1024x1024, 512 markers, MDF(5) + Guo forcing, "
                                       "2-DOF Newmark body on device" + (f"; per GPU, {world} slabs per channel" if world > 1 else ""),
                           "cells_per_step": cells, "ensemble_domains": n_rep,
                           "l2": f"{n_rep} independent domains rotated so the working set ({n_rep * bytes_per_domain / 1e6:.0f} MB) "
                                 "exceeds 4x L2: inputs come from HBM every step (no L2 flush needed)",
                           "l2_resident_mlups": l2_mlups,
                           "hbm_frac_of_measured": value / world * 1e6 * 72 / (hbm_gbs * 1e9),
                           "multi_gpu": multi,
                           "state_finite": finite},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": inner[0].n_launch_per_step * K,
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
    ap.add_argument("--steps", type=int, default=50000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the C3 / C4 single-GPU measurements")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
