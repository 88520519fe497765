"""Where does the C5 step (1024 x 512 x 512 MRT, 695 k markers) go?  Graph-replayed steps with parts left out (timing
only):  python scripts/c5_parts.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vivsim_b200 import Stepper, configs, _lib as L
import vivsim_b200.stepper as S

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
        if k in ("vsb_ib_mdf",):
            return (lambda *a: 0) if "mdf" in skip else fn
        if k in ("vsb_ib_window_moments", "vsb_ib_window_moments_cells"):
            return (lambda *a: 0) if "uwin" in skip else fn
        return fn


S.L.lib = lambda: Proxy()
small = "small" in sys.argv
if "c3" in sys.argv:
    spec, body = configs.sphere_3d()
else:
    spec, body = configs.oscillating_cylinder_3d(nx=256, ny=256, nz=256) if small else configs.oscillating_cylinder_3d()
cells = bench.cells_of(spec)
f0 = configs.uniform_state(spec, noise=1e-3)
st = Stepper(spec, body=dict(body), dyn_mode="device", follow=2) if body else Stepper(spec)
st.set_f(f0); st.step(3)
del f0
print("reach list:", None if st._reach is None else st._reach.numel(), "window", st.win_size, flush=True)
for what in ([], ["uwin"], ["uwin", "mdf"], ["uwin", "mdf", "band"], ["bulk"], ["bulk", "band"], ["bulk", "band", "mdf"],
             ["bulk", "uwin", "mdf"]):
    skip.clear(); skip.update(what)
    loop = bench.GraphLoop([st], 2)
    loop.run(1)
    n = 4 if "c3" not in sys.argv else 20
    dt, _, _ = bench.timed(lambda: loop.run(n), torch.cuda.synchronize)
    ran = [p for p in ("uwin", "mdf", "band", "bulk") if p not in skip]
    ms = dt / (n * 2) * 1e3
    print(f"runs {'+'.join(ran):22s}: {ms:7.4f} ms per step  ({cells * 152 / ms / 1e6 / 6451.5:5.3f} of HBM if it were the whole step)", flush=True)
    del loop
skip.clear()
plain = dict(spec); plain.pop("ib")
del st
torch.cuda.empty_cache()
st = Stepper(plain); st.set_f(configs.uniform_state(plain, noise=1e-3)); st.step(3)
loop = bench.GraphLoop([st], 2); loop.run(1)
dt, _, _ = bench.timed(lambda: loop.run(4), torch.cuda.synchronize)
ms = dt / 8 * 1e3
print(f"no body (walls only)        : {ms:7.3f} ms per step  ({cells * 152 / ms / 1e6 / 6451.5:5.3f} of HBM)")
