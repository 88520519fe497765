"""Multi-GPU parity worker (run under torchrun on N GPUs of one node; tests/test_gpu_multi.py launches it):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_worker.py

Every rank runs its slab of (a) a periodic D2Q9 KBC + EDM body-force case, (b) the C2 recipe with walls and one
immersed cylinder per slab, (c) a D3Q19 BGK case, (d) a D3Q19 MRT case with walls and a densely meshed immersed cylinder; rank 0 also runs the whole domain on one GPU and compares:
bit-exact for every case (same per-cell arithmetic, exchange is a copy)."""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vivsim_b200 import Stepper, configs  # noqa: E402
from vivsim_b200.multidevice import SlabStepper  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    ok = True

    def compare(name, spec, f0, n, local_ib=None, global_ib_spec=None):
        for halo in ("nccl", "peer"):
            _compare(f"{name} / {halo}", spec, f0, n, halo, local_ib, global_ib_spec)

    def _compare(name, spec, f0, n, halo, local_ib, global_ib_spec):
        nonlocal ok
        s = SlabStepper(spec, local_ib=local_ib, halo=halo).set_f_global(f0)
        s.step(n // 2)
        if halo == "peer":           # second half through a captured CUDA graph (2 steps per replay)
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                s.advance_raw(2)
            torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize(); dist.barrier()
            with torch.cuda.graph(g):
                s.advance_raw(2)
            for _ in range((n - n // 2 - 2) // 2):
                g.replay()
            torch.cuda.synchronize()
            assert not s.peer.timed_out(), "halo wait timed out"
        else:
            s.step(n - n // 2)
        got = s.gather_f()
        force = s.total_force()
        if rank == 0:
            ref_spec = global_ib_spec if global_ib_spec is not None else spec
            refs = Stepper(ref_spec).set_f(f0)
            refs.step(n)
            ref = refs.get_f()
            same = torch.equal(got, ref)
            err = float((got - ref).abs().max() / ref.abs().max())
            print(f"[{name}] world={world} bit-exact={same} max rel diff={err:.2e} total force={force.tolist()}")
            ok = ok and (same or err < (1e-5 if "ib" in name and spec["dim"] == 3 else 1e-6))

    # (a) periodic KBC + uniform force
    shape = (64 * world, 96)
    spec = dict(dim=2, shape=shape, collision="kbc", omega=1.8, forcing="edm", g=(1e-5, -2e-5), post=[])
    gen = torch.Generator(device="cuda").manual_seed(0)
    f0 = configs.uniform_state(dict(spec, u0=0.05), noise=1e-3)
    dist.broadcast(f0, 0)
    compare("periodic kbc", spec, f0, 20)

    # (b) C2 recipe with walls; one cylinder in the slab of the last rank (global spec has a single body)
    nx, ny = 128 * world, 128
    spec, _ = configs.viv_cylinder_2d(nx=nx, ny=ny, n_marker=64, radius=8.0, u0=0.08, nu=0.02, moving=False,
                                      center=(nx - 64.0, ny / 2))
    f0 = configs.uniform_state(spec, noise=1e-3)
    dist.broadcast(f0, 0)
    compare("c2 walls + ib", spec, f0, 20)

    # (c) D3Q19 BGK periodic
    shape = (16 * world, 12, 32)
    spec = dict(dim=3, shape=shape, collision="bgk", omega=1.6, forcing="guo", g=(1e-5, 0.0, 2e-5), post=[])
    f0 = configs.uniform_state(dict(spec, u0=0.04), noise=1e-3)
    dist.broadcast(f0, 0)
    compare("3d bgk", spec, f0, 12)

    # (d) D3Q19 MRT + Guo with walls and a densely meshed fixed cylinder (tiled MDF, 3-D force-window band) in the
    #     slab of rank 0; spreading uses fp32 atomics, so the comparison is to rounding, not bit-exact
    nx = 64 * world
    spec, _ = configs.oscillating_cylinder_3d(nx=nx, ny=48, nz=48, diameter=12.0, center_x=30.0, moving=False)
    f0 = configs.uniform_state(spec, noise=1e-3)
    dist.broadcast(f0, 0)
    compare("3d mrt walls + dense ib", spec, f0, 12)

    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL")
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
