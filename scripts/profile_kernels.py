"""Run a few fused steps of one configuration (for ncu):  python scripts/profile_kernels.py c3|c4|c2|c5 [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vivsim_b200 import Stepper, configs  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
if name == "c2":
    spec, body = configs.viv_cylinder_2d()
elif name == "c3":
    spec, body = configs.sphere_3d()
elif name == "c4":
    spec, body = configs.viv_cylinder_2d_large(n=8192)
elif name == "c5":   # the C5 recipe (MRT + Guo-MRT, NEBB / equilibrium faces, moving finite cylinder, tiled MDF) at 256^3
    spec, body = configs.oscillating_cylinder_3d(nx=256, ny=256, nz=256)
elif name == "c5slab":   # one slab of the 1024 x 512 x 512 MRT case (1/8 of the domain), periodic, uniform Guo body force
    spec, body = dict(dim=3, shape=(128, 512, 512), collision="mrt", omega=1.7, forcing="guo", g=(1e-6, 0.0, 0.0),
                      post=[], u0=0.05), None
else:
    raise SystemExit("unknown workload")
kw = dict(ib_chain=os.environ["VSB_CHAIN"]) if "VSB_CHAIN" in os.environ else {}
st = Stepper(spec, body=body, dyn_mode="device", follow=2 if name == "c5" else 1, **kw) if body else Stepper(spec, **kw)
st.set_f(configs.uniform_state(spec, noise=1e-3))
st.step(steps)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); st.step(steps); e1.record(); torch.cuda.synchronize()
cells = 1
for n in spec["shape"]:
    cells *= n
print(name, "ms/step", e0.elapsed_time(e1) / steps, "MLUPS", cells * steps / e0.elapsed_time(e1) / 1e3)
