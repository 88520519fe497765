"""C2 ensemble throughput vs. number of domains and steps per graph replay."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vivsim_b200 import Stepper, configs

spec, body = configs.viv_cylinder_2d()
f0 = configs.uniform_state(spec, noise=1e-3)
for n_rep, steps_each in ((8, 4), (8, 8), (8, 16), (7, 8), (6, 8), (4, 8)):
    sts = []
    for _ in range(n_rep):
        st = Stepper(spec, body=dict(body), dyn_mode="device").set_f(f0); st.step(1); sts.append(st)
    g = bench.build_graph(sts, steps_each)
    K = n_rep * steps_each * 40
    bench.run_loop(g, sts, steps_each, n_rep * steps_each * 5)
    dt, _, _ = bench.timed(lambda: bench.run_loop(g, sts, steps_each, K), torch.cuda.synchronize)
    print(f"domains {n_rep:3d} steps/graph/domain {steps_each}: {dt / K * 1e6:6.2f} us/step  {1048576 * K / dt / 1e9:6.2f} GLUPS")
    del sts, g
    torch.cuda.empty_cache()
