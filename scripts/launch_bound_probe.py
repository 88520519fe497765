"""Is the C2 ensemble step bound by kernel count?  Time it for several MDF iteration counts and without IB."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vivsim_b200 import Stepper, configs

def run(tag, make, n_rep=8, K=4000):
    sts = []
    for _ in range(n_rep):
        st = make(); sts.append(st)
    g = bench.build_graph(sts, 2)
    bench.run_loop(g, sts, 2, 200)
    dt, _, _ = bench.timed(lambda: bench.run_loop(g, sts, 2, K), torch.cuda.synchronize)
    print(f"{tag:34s} {dt / K * 1e6:7.2f} us/step  launches/step {sts[0].n_launch_per_step}")

spec, body = configs.viv_cylinder_2d()
f0 = configs.uniform_state(spec, noise=1e-3)
def mk(n_iter=5, ib=True, post=True, **kw):
    def f():
        sp = dict(spec)
        if not ib: sp.pop("ib")
        else: sp["ib"] = dict(spec["ib"], n_iter=n_iter)
        if not post: sp["post"] = []
        st = Stepper(sp, body=dict(body) if ib else None, dyn_mode="device", **kw).set_f(f0); st.step(1); return st
    return f
run("no IB, no walls (1 launch)", mk(ib=False, post=False))
run("no IB, walls", mk(ib=False))
for n in (1, 2, 3, 5):
    run(f"IB n_iter={n}, walls", mk(n_iter=n))
run("IB n_iter=5, walls, no overlap", mk(n_iter=5, overlap=False))
