"""BASELINE configs 3 and 4 on N GPUs of one node (strong scaling of a fixed global grid):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29540 \
        scripts/config_runs.py --config c4|c5 [--scale 1.0] [--steps 20]

c4: D2Q9 KBC VIV cylinder Re = 1e4 on 16384 x 16384, slab-decomposed.
c5: D3Q19 MRT elastically mounted cylinder on 1024 x 512 x 512 with IB, slab-decomposed along x (a slab of
    1024/8 x 512 x 512 has 1.6 % surface-to-volume, so x-slabs are used instead of pencils).
Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vivsim_b200 import Stepper, configs  # noqa: E402
from vivsim_b200.multidevice import SlabStepper  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c4")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    if args.config == "c4":
        n = int(16384 * args.scale)
        spec, body = configs.viv_cylinder_2d_large(n=n)
        bpc, label = 72, f"C4 D2Q9 KBC VIV cylinder {n}x{n}, {spec['ib']['markers'].shape[0]} markers, MDF(5) + EDM"
    else:
        nx, ny, nz = int(1024 * args.scale), int(512 * args.scale), int(512 * args.scale)
        spec, body = configs.oscillating_cylinder_3d(nx=nx, ny=ny, nz=nz)
        bpc = 152
        label = f"C5 D3Q19 MRT oscillating cylinder {nx}x{ny}x{nz}, {spec['ib']['markers'].shape[0]} markers, MDF(3) + Guo-MRT"
    cells = 1
    for k in spec["shape"]:
        cells *= k
    if world > 1:
        st = SlabStepper(spec, body=body, dyn_mode="device", follow=2 if spec["dim"] == 3 else 1)
        slab = st.slab
        lshape = slab.local_shape
        loc = dict(spec, shape=lshape)
        st.set_f_local(configs.uniform_state(loc, noise=1e-3))
        halo = st.halo
        inner = st.stepper
    else:
        st = Stepper(spec, body=body, dyn_mode="device", follow=2 if spec["dim"] == 3 else 1)
        st.set_f(configs.uniform_state(spec, noise=1e-3))
        halo, inner = "none", st
    st.step(3)
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        st.advance_raw(2)
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    use_graph = halo in ("peer", "none")
    if use_graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            st.advance_raw(2)
        run = lambda: [g.replay() for _ in range(args.steps // 2)]
    else:
        run = lambda: st.step(args.steps)
    run(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3
    if world > 1:
        t = torch.tensor([dt], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t)
    steps = (args.steps // 2) * 2 if use_graph else args.steps
    finite = torch.tensor([1.0 if bool(torch.isfinite(inner.state).all()) else 0.0], device="cuda")
    if world > 1:
        dist.all_reduce(finite, op=dist.ReduceOp.MIN)
    if world > 1 and st.peer is not None:
        assert not st.peer.timed_out(), "halo wait timed out"
    if rank == 0:
        mlups = cells * steps / dt / 1e6
        hbm = 6452.8
        try:
            hbm = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        print(json.dumps({"config": label, "n_gpus": world, "halo": halo, "steps": steps, "ms_per_step": dt / steps * 1e3,
                          "mlups": mlups, "per_gpu_frac_of_measured_hbm_roofline": mlups / world * 1e6 * bpc / (hbm * 1e9),
                          "finite": bool(finite.item()), "launches_per_step": st.n_launch_per_step}))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
