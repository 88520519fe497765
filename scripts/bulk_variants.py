"""What makes the bulk pass of the C2 step (band = 1, 8-domain ensemble) slower than the plain fused kernel?  Timing
only (chain and band are skipped, so the physics is wrong):  python scripts/bulk_variants.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vivsim_b200 import Stepper, configs, _lib as L
import vivsim_b200.stepper as S

real_lib = L.lib()


class Proxy:
    def __getattr__(self, k):
        fn = getattr(real_lib, k)
        if k == "vsb_step":
            return lambda ref, stm: 0 if ref._obj.band == 2 else fn(ref, stm)
        if k == "vsb_ib_mdf":
            return lambda *a: 0
        return fn


S.L.lib = lambda: Proxy()
spec, body = configs.viv_cylinder_2d()
f0 = configs.uniform_state(spec, noise=1e-3)
plain = dict(spec); plain.pop("ib"); plain["post"] = []
nowall = dict(spec, post=[])
cases = [("plain kernel (no IB, no walls)", plain, None),
         ("walls only (no IB)", dict(plain, post=spec["post"]), None),
         ("IB bulk, moving body, walls", spec, body),
         ("IB bulk, moving body, no walls", nowall, body),
         ("IB bulk, fixed body, walls", spec, None),
         ("IB bulk, fixed body, no walls", nowall, None)]
for name, sp, bd in cases:
    sts = []
    for _ in range(8):
        st = Stepper(sp, body=dict(bd), dyn_mode="device") if bd else Stepper(sp)
        st.set_f(f0); st.step(1)
        sts.append(st)
    loop = bench.GraphLoop(sts, 10)
    loop.run(20)
    n = 100
    dt, _, _ = bench.timed(lambda: loop.run(n), torch.cuda.synchronize)
    print(f"{name:34s}: {dt / (n * 10 * 8) * 1e6:6.2f} us per step per domain", flush=True)
    del loop, sts
