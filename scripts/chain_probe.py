"""C2 step time for every way of chaining the MDF iterations (ib_chain): one L2-resident domain and the 8-domain
ensemble of bench.py.  python scripts/chain_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vivsim_b200 import Stepper, configs

spec, body = configs.viv_cylinder_2d()
cells = bench.cells_of(spec)
f0 = configs.uniform_state(spec, noise=1e-3)
sync = torch.cuda.synchronize
for n_dom in (1, 8):
    for chain, first in (("cluster", False), ("cluster", True), ("barrier", False), ("barrier", True), ("launches", False)):
        sts = []
        for _ in range(n_dom):
            st = Stepper(spec, body=dict(body), dyn_mode="device", ib_chain=chain, chain_first=first)
            st.set_f(f0); st.step(1)
            sts.append(st)
        loop = bench.GraphLoop(sts, 10)
        loop.run(20)
        n = 200
        dt, _, _ = bench.timed(lambda: loop.run(n), sync)
        us = dt / (n * 10 * n_dom) * 1e6
        d, v, a, h = sts[0].body_state()
        print(f"domains {n_dom} chain {chain:9s} chain_first {int(first)}: {us:7.2f} us per lattice step per domain, {cells / us / 1e3:8.1f} GLUPS, "
              f"d = {d}, h = {h}", flush=True)
        del loop, sts
        torch.cuda.empty_cache()
