"""Where does the period of ONE L2-resident C2 domain go?  Graph-replayed steps with parts of the step left out (timing
only -- the physics is wrong without them):  python scripts/single_domain_parts.py [chain ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vivsim_b200 import Stepper, configs, _lib as L

spec, body = configs.viv_cylinder_2d()
cells = bench.cells_of(spec)
f0 = configs.uniform_state(spec, noise=1e-3)
sync = torch.cuda.synchronize
real_lib = L.lib()
skip = set()


class Proxy:
    def __getattr__(self, k):
        fn = getattr(real_lib, k)
        if k == "vsb_step":
            def step(ref, stm):
                band = ref._obj.band
                if ("bulk" in skip and band == 1) or ("band" in skip and band == 2):
                    return 0
                return fn(ref, stm)
            return step
        if k == "vsb_ib_mdf":
            def mdf(*a):
                return 0 if "chain" in skip else fn(*a)
            return mdf
        return fn


L.lib = lambda: Proxy()
import vivsim_b200.stepper as S
S.L.lib = L.lib
chains = sys.argv[1:] or ["cluster", "barrier"]
for chain in chains:
    for first in (False, True):
        combos = ([], ["bulk"], ["bulk", "band"], ["bulk", "chain"], ["chain", "band"], ["chain"], ["band"])
        if os.environ.get("VSB_PARTS_ONLY") == "chain":
            combos = (["bulk", "band"],)
            if first:
                continue
        for what in combos:
            skip.clear(); skip.update(what)
            n_dom = int(os.environ.get("VSB_PARTS_DOMAINS", "1"))
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
            ran = [p for p in ("chain", "band", "bulk") if p not in skip]
            print(f"domains {n_dom} chain {chain:8s} chain_first {int(first)}  runs {'+'.join(ran):16s}: {us:6.2f} us per step", flush=True)
            del loop, st, sts
