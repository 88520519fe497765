"""C3 with the IB chain switched off (force window stays zero): what the fused kernel costs with the window logic and
the band split alone.   python scripts/c3_nochain.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vivsim_b200 import Stepper, configs

base, _ = configs.sphere_3d()
cells = 256 ** 3


def run(tag, spec, steps=20, chain=True, **kw):
    st = Stepper(spec, **kw).set_f(configs.uniform_state(spec, noise=1e-3)); st.step(4)
    if not chain:
        st._ib_part = lambda stream: None
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); st.step(steps); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{tag:58s} {ms * 1e3:8.1f} us/step  {cells * 152 / ms / 1e6 / 6553.3:5.3f} of HBM", flush=True)


body = dict(base, post=[])
run("periodic, no forcing", dict(base, post=[], ib=None, forcing=None))
run("body, chain on, overlap on", body)
run("body, chain OFF, overlap on (band 2 + band 1 launches)", body, chain=False)
run("body, chain OFF, overlap off (one band-0 launch)", body, chain=False, overlap=False)
run("body, chain on, overlap off", body, overlap=False)
