"""Phases of the one-CTA MDF chain (VSB_CTA_STOP=1..6, 0 = everything) on the C2 cylinder held FIXED (no body ODE, so
that leaving parts of the step out does not change the marker positions):  the chain alone in a graph replay, and the
whole step.  python scripts/cta_chain_phases.py [cta|barrier]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vivsim_b200 import Stepper, configs, _lib as L

spec, body = configs.viv_cylinder_2d()
f0 = configs.uniform_state(spec, noise=1e-3)
real_lib = L.lib()
skip = set()


class Proxy:
    def __getattr__(self, k):
        fn = getattr(real_lib, k)
        if k == "vsb_step":
            return (lambda ref, stm: 0) if "fluid" in skip else fn
        return fn


import vivsim_b200.stepper as S
S.L.lib = lambda: Proxy()
chain = sys.argv[1] if len(sys.argv) > 1 else "cta"
for what in (["fluid"], []):
    skip.clear(); skip.update(what)
    st = Stepper(spec, ib_chain=chain)          # static body
    st.set_f(f0); st.step(1)
    loop = bench.GraphLoop([st], 10)
    loop.run(20)
    n = 200
    dt, _, _ = bench.timed(lambda: loop.run(n), torch.cuda.synchronize)
    print(f"chain {chain} stop {os.environ.get('VSB_CTA_STOP', '0')}  {'chain alone' if what else 'whole step '}: {dt / (n * 10) * 1e6:7.2f} us per step", flush=True)
    del loop, st
