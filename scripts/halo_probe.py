"""One GPU: what the halo pipeline costs the C2 ensemble step.  The same 8-domain ensemble (bench.py's `value`) as
plain steppers, as one-rank slabs through the three-launch pipeline (wait / edge rows / send; local periodic halo) and
through the fused edge-row launch (VsbStepArgs.halo).  On one GPU the flag waits return at once, so the difference to
the plain steppers is launch structure only -- the part of the weak-scaling loss that is not NVLink latency."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from vivsim_b200 import Stepper, configs  # noqa: E402
from vivsim_b200.multidevice import SlabStepper  # noqa: E402

spec, body = configs.viv_cylinder_2d()
f0 = configs.uniform_state(spec, noise=1e-3)
cells = bench.cells_of(spec)
n_rep, steps_each, replays = 8, bench.GRAPH_STEPS, 300
f_loc = torch.cat([f0[:, -1:], f0, f0[:, :1]], dim=1).contiguous()
gspec = dict(spec)
gspec.pop("ib")


def local_ib(slab):
    return spec["ib"]


for mode in ("plain", "pipelined-local", "fused-local"):
    sts = []
    for _ in range(n_rep):
        if mode == "plain":
            st = Stepper(spec, body=dict(body), dyn_mode="device", ib_chain=bench.ENSEMBLE_CHAIN)
            st.set_f(f0)
        else:
            st = SlabStepper(gspec, rank=0, world=1, local_ib=local_ib, body=dict(body), dyn_mode="device", halo=mode,
                             ib_chain=bench.ENSEMBLE_CHAIN)
            st.set_f_local(f_loc)
        st.step(1)
        sts.append(st)
    loop = bench.GraphLoop(sts, steps_each)
    loop.run(50)
    dt, _, _ = bench.timed(lambda: loop.run(replays), torch.cuda.synchronize)
    us = dt / (replays * steps_each * n_rep) * 1e6
    inner = sts[0] if mode == "plain" else sts[0].stepper
    print(f"{mode:16s} {us:7.2f} us per lattice step  {cells / us / 1e3:7.1f} GLUPS  launches/step {inner.n_launch_per_step}", flush=True)
    del loop, sts
    torch.cuda.empty_cache()
