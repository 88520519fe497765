"""Where the C3 step (D3Q19 KBC sphere 256^3) spends its time: the same grid with and without walls / immersed body
and with the scheduling options switched off one at a time.   python scripts/c3_breakdown.py [n]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vivsim_b200 import Stepper, configs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
base, _ = configs.sphere_3d(nx=n, ny=n, nz=n, diameter=48.0 * n / 256)
cells = n ** 3


def run(tag, spec, steps=20, **kw):
    st = Stepper(spec, **kw).set_f(configs.uniform_state(spec, noise=1e-3)); st.step(4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); st.step(steps); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{tag:58s} {ms * 1e3:8.1f} us/step {cells / ms / 1e3:8.0f} MLUPS  {cells * 152 / ms / 1e6 / 6452.8:5.3f} of HBM  "
          f"launches/step {st.n_launch_per_step}", flush=True)
    del st


periodic = dict(base, post=[], ib=None, forcing=None)
run("periodic, no forcing (vec_sweep case)", periodic)
run("periodic, forcing=edm uniform g", dict(base, post=[], ib=None, g=(1e-6, 0.0, 0.0)))
run("walls only (nebb left / equilibrium right), no body", dict(base, ib=None, forcing=None))
run("walls only, fuse_edges off", dict(base, ib=None, forcing=None), fuse_edges=False)
run("body only (periodic x)", dict(base, post=[]))
run("body only, overlap off", dict(base, post=[]), overlap=False)
run("full C3", base)
run("full C3, CUDA graph", base, use_graph=True)
run("full C3, overlap off", base, overlap=False)
run("full C3, vec=4", base, vec=4)
run("full C3, vec=1", base, vec=1)
