"""When do the branches of one overlapped step finish?  Records CUDA events after the IB chain, after the window-band
launch and after the bulk launch and prints their offsets from the fork.   python scripts/step_timeline.py [c3|c2|c5]"""
import os, sys
import ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vivsim_b200 import Stepper, configs, _lib as L

which = sys.argv[1] if len(sys.argv) > 1 else "c3"
if which == "c3":
    spec, _ = configs.sphere_3d(nx=256, ny=256, nz=256, diameter=48.0)
elif which == "c2":
    spec, _ = configs.viv_cylinder_2d()
else:
    raise SystemExit("c3 | c2")
st = Stepper(spec).set_f(configs.uniform_state(spec, noise=1e-3)); st.step(4)
torch.cuda.synchronize()
lib = L.lib()
marks = []
orig_step, orig_ib = lib.vsb_step, Stepper._ib_part


def rec(tag, stream_ptr):
    ev = torch.cuda.Event(enable_timing=True)
    ev.record(torch.cuda.ExternalStream(stream_ptr) if stream_ptr else torch.cuda.default_stream())
    marks.append((tag, ev))


def step_wrap(ref, stm):
    r = orig_step(ref, stm)
    rec("k_step band=%d" % st._args.band, stm.value)
    return r


def ib_wrap(self, stm):
    orig_ib(self, stm)
    rec("ib chain", stm.value)


class LibProxy:
    def __getattr__(self, k):
        return step_wrap if k == "vsb_step" else getattr(lib, k)


L_lib = L.lib
L.lib = lambda: LibProxy()
Stepper._ib_part = ib_wrap
acc = {}
n = 20
for it in range(n):
    marks.clear()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e0.record()
    st._advance()
    e1 = torch.cuda.Event(enable_timing=True); e1.record()
    torch.cuda.synchronize()
    for tag, ev in marks + [("join", e1)]:
        acc.setdefault(tag, []).append(e0.elapsed_time(ev) * 1e3)
for tag, v in acc.items():
    v = sorted(v)
    print(f"{which} {tag:18s} done at {v[len(v) // 2]:8.1f} us (median of {n}, single isolated step, VSB_IB_PRIORITY={os.environ.get('VSB_IB_PRIORITY', '-1')})")
