"""IB chain beside the bulk pass (overlap) or before one pass over the whole grid (no overlap)?  C3, C5, C4, C2:
    python scripts/overlap_probe.py [c3 c5 c4 c2]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vivsim_b200 import Stepper, configs

which = [a for a in sys.argv[1:]] or ["c3", "c5", "c4", "c2"]
for name in which:
    if name == "c3":
        spec, body, bpc, kw = *configs.sphere_3d(), 152, {}
    elif name == "c5":
        spec, body, bpc, kw = *configs.oscillating_cylinder_3d(), 152, dict(follow=2)
    elif name == "c4":
        spec, body, bpc, kw = *configs.viv_cylinder_2d_large(), 72, {}
    else:
        spec, body, bpc, kw = *configs.viv_cylinder_2d(), 72, {}
    cells = bench.cells_of(spec)
    f0 = configs.uniform_state(spec, noise=1e-3)
    for overlap in (True, False):
        st = Stepper(spec, body=dict(body), dyn_mode="device", overlap=overlap, **kw) if body else Stepper(spec, overlap=overlap)
        st.set_f(f0); st.step(3)
        loop = bench.GraphLoop([st], 2)
        loop.run(2)
        n = 6 if name in ("c5", "c4") else 50
        dt, _, _ = bench.timed(lambda: loop.run(n), torch.cuda.synchronize)
        ms = dt / (2 * n) * 1e3
        print(f"{name} overlap {int(overlap)}: {ms:8.4f} ms per step  {cells / ms / 1e6:7.2f} GLUPS  {cells * bpc / ms / 1e6 / 6451.5:5.3f} of HBM", flush=True)
        del loop, st
        torch.cuda.empty_cache()
    del f0
    torch.cuda.empty_cache()
