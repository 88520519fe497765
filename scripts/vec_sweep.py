"""Time the fused kernel alone (no IB, no walls) for several vector widths:  python scripts/vec_sweep.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vivsim_b200 import Stepper, configs

def run(tag, spec, bpc, vecs=(1, 2, 4), steps=20):
    cells = 1
    for n in spec["shape"]: cells *= n
    f0 = configs.uniform_state(spec, noise=1e-3)
    for vec in vecs:
        st = Stepper(spec, vec=vec).set_f(f0); st.step(3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); st.advance_raw(steps); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(f"{tag:28s} vec={vec} {ms*1e3:9.1f} us/step  {cells/ms/1e3:9.0f} MLUPS  {cells*bpc/ms/1e6/6452.8:5.3f} of HBM")
        del st

if "quick3d" in sys.argv:
    for coll in ("bgk", "kbc", "mrt"):
        run(f"D3Q19 {coll} 256^3", dict(dim=3, shape=(256, 256, 256), collision=coll, omega=1.7, forcing=None, post=[], u0=0.05), 152, vecs=(2, 4))
    run("D3Q19 mrt+guo(uniform) 256^3", dict(dim=3, shape=(256, 256, 256), collision="mrt", omega=1.7, forcing="guo", g=(1e-6, 0.0, 0.0), post=[], u0=0.05), 152, vecs=(2, 4))
    sys.exit(0)
if "quick" in sys.argv:
    for coll in ("bgk", "kbc", "mrt"):
        run(f"D3Q19 {coll} 256^3", dict(dim=3, shape=(256, 256, 256), collision=coll, omega=1.7, forcing=None, post=[], u0=0.05), 152, vecs=(2, 4))
    for coll in ("bgk", "kbc"):
        run(f"D2Q9 {coll} 8192^2", dict(dim=2, shape=(8192, 8192), collision=coll, omega=1.7, forcing=None, post=[], u0=0.05), 72, vecs=(4,), steps=8)
    sys.exit(0)
for coll in ("bgk", "kbc", "reg", "mrt"):
    run(f"D3Q19 {coll} 256^3", dict(dim=3, shape=(256, 256, 256), collision=coll, omega=1.7, forcing=None, post=[], u0=0.05), 152)
if "3d" in sys.argv: sys.exit(0)
for coll in ("bgk", "kbc", "reg", "mrt"):
    run(f"D2Q9 {coll} 8192^2", dict(dim=2, shape=(8192, 8192), collision=coll, omega=1.7, forcing=None, post=[], u0=0.05), 72, steps=8)
