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

    def compare(name, spec, f0, n, local_ib=None, global_ib_spec=None, halos=("nccl", "peer"), **kw):
        for halo in halos:
            _compare(f"{name} / {halo}", spec, f0, n, halo, local_ib, global_ib_spec, **kw)

    def _compare(name, spec, f0, n, halo, local_ib, global_ib_spec, ib="auto", body=None, follow=1, tol=None):
        nonlocal ok
        extra = dict(body=dict(body), dyn_mode="device", follow=follow) if body is not None else {}
        s = SlabStepper(spec, local_ib=local_ib, halo=halo, ib=ib, **extra).set_f_global(f0)
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
        mf = s.marker_force() if s.ib_mode == "shard" else None
        if s.ib_shard is not None:
            assert not s.ib_shard.timed_out(), "a barrier of the shared IB chain timed out"
        if rank == 0:
            ref_spec = global_ib_spec if global_ib_spec is not None else spec
            refs = Stepper(ref_spec, **extra).set_f(f0)
            refs.step(n)
            ref = refs.get_f()
            same = torch.equal(got, ref)
            err = float((got - ref).abs().max() / ref.abs().max())
            line = f"[{name}] world={world} ib={s.ib_mode} bit-exact={same} max rel diff={err:.2e} total force={force.tolist()}"
            bound = tol if tol is not None else (1e-5 if "ib" in name and spec["dim"] == 3 else 1e-6)
            good = same or err < bound
            if mf is not None:      # shared chain: every marker's force, gathered from the ranks' shares
                rf = refs.marker_force
                ferr = float((mf - rf).abs().max() / rf.abs().max())
                line += f" marker force rel diff={ferr:.2e}"
                good = good and ferr < 2e-4
            if body is not None:    # replicas of the rigid body on every rank vs the single-GPU body
                mine = np.concatenate(s.stepper.body_state())
                theirs = np.concatenate(refs.body_state())
                berr = float(np.abs(mine - theirs).max() / max(np.abs(theirs).max(), 1e-30))
                line += f" body rel diff={berr:.2e}"
                good = good and berr < 1e-4
            print(line, flush=True)
            ok = ok and good
        if body is not None:        # every replica must be IDENTICAL (same sums in the same order)
            st = torch.as_tensor(np.concatenate(s.stepper.body_state()), device="cuda")
            lo, hi = st.clone(), st.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            if not torch.equal(lo, hi):
                print(f"[{name}] rank {rank}: body replicas differ between ranks", flush=True)
                ok = False

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

    # ---- immersed-boundary chain shared by all ranks (ib='shard'): bodies ON a cut, moving across it, rotating
    # (e) C2 recipe, fixed cylinder centred exactly on the cut between the last two slabs
    nx, ny = 128 * world, 128
    spec, _ = configs.viv_cylinder_2d(nx=nx, ny=ny, n_marker=64, radius=8.0, u0=0.08, nu=0.02, moving=False,
                                      center=(nx - 128.0, ny / 2))
    f0 = configs.uniform_state(spec, noise=1e-3)
    dist.broadcast(f0, 0)
    compare("c2 ib on a cut", spec, f0, 20, ib="shard", tol=1e-5)

    # (f) elastically mounted cylinder (2-DOF Newmark on every rank's replica) starting on a cut and carried by the flow
    spec, body = configs.viv_cylinder_2d(nx=nx, ny=ny, n_marker=64, radius=8.0, u0=0.08, nu=0.02, moving=True,
                                         center=(nx - 128.0 - 0.4, ny / 2))
    body = dict(body, v0=(0.02, 0.01), k=0.0)          # free body with an initial velocity: it crosses cells and the cut
    f0 = configs.uniform_state(spec, noise=1e-3)
    dist.broadcast(f0, 0)
    compare("c2 moving ib across a cut", spec, f0, 40, ib="shard", body=body, tol=1e-5)

    # (g) rotating ellipse (3 degrees of freedom: torque all-reduce) on a cut
    th = np.linspace(0, 2 * np.pi, 48, endpoint=False)
    cx, cy = nx - 128.0 + 0.3, ny / 2 + 0.2
    mk = np.stack([cx + 9.0 * np.cos(th), cy + 5.0 * np.sin(th)], axis=1).astype(np.float32)
    spec_r = dict(spec)
    from vivsim_b200 import ib as ib2
    spec_r["ib"] = dict(markers=mk, ds=ib2.get_ds(mk), kernel="peskin4", n_iter=3, u_target=None,
                        window=((int(cx) - 16, int(cy) - 16), (32, 32)))
    area = np.pi * 45.0
    body_r = dict(m=np.diag([10 * area, 10 * area, 2.5 * area * 106.0]), k=np.diag([0.02, 0.05, 0.9]),
                  c=np.diag([0.01, 0.02, 0.3]), added_mass=np.array([area, area, 0.0]), n_dof=3, rotation=True,
                  center=(cx, cy), d0=(0.0, 0.0, 0.3), v0=(0.0, 0.01, 0.004), a0=(0.0, 0.0, 0.0))
    compare("c2 rotating ib on a cut", spec_r, f0, 32, ib="shard", body=body_r, tol=1e-5)

    # (h) D3Q19 MRT with the densely meshed cylinder of (d): the tiled chain divided along the cylinder's axis
    nx = 64 * world
    spec, _ = configs.oscillating_cylinder_3d(nx=nx, ny=48, nz=48, diameter=12.0, center_x=30.0, moving=False)
    f0 = configs.uniform_state(spec, noise=1e-3)
    dist.broadcast(f0, 0)
    compare("3d mrt walls + dense ib shared", spec, f0, 12, ib="shard")
    spec, body3 = configs.oscillating_cylinder_3d(nx=nx, ny=48, nz=48, diameter=12.0, center_x=62.5, moving=True)
    f0 = configs.uniform_state(spec, noise=1e-3)
    dist.broadcast(f0, 0)
    compare("3d mrt walls + dense moving ib on a cut", spec, f0, 12, ib="shard", body=body3, follow=2, halos=("peer",))

    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL")
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
