"""C5 (1024 x 512 x 512 MRT, 695 k markers): the IB chain alone and the whole step, for the library / tile geometry the
environment selects (VIVSIM_B200_LIB, VSB_TILE_CELLS, VSB_TILE_COLUMN).  python scripts/tiled_probe.py [small]"""
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
            return (lambda ref, stm: 0) if "fluid" in skip else fn
        return fn


S.L.lib = lambda: Proxy()
small = "small" in sys.argv
spec, body = configs.oscillating_cylinder_3d(nx=256, ny=256, nz=256) if small else configs.oscillating_cylinder_3d()
cells = bench.cells_of(spec)
f0 = configs.uniform_state(spec, noise=1e-3)
st = Stepper(spec, body=dict(body), dyn_mode="device", follow=2)
st.set_f(f0); st.step(3)
del f0
tag = f"lib={os.path.basename(L.LIB_PATH)} cells={S.TILE_CELLS} column={S.TILE_COLUMN} chunks={st._chunk_offsets.numel() - 1}"
for what in (["fluid"], []):
    skip.clear(); skip.update(what)
    loop = bench.GraphLoop([st], 2)
    loop.run(2)
    n = 6
    dt, _, _ = bench.timed(lambda: loop.run(n), torch.cuda.synchronize)
    print(f"{tag}  {'chain alone' if what else 'whole step '}: {dt / (2 * n) * 1e3:8.4f} ms per step", flush=True)
    del loop
