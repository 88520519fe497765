"""Reproduce the e2e leg of bench.py in isolation (debugging aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vivsim_b200 import Stepper, Ensemble, configs

hist = int(sys.argv[1]) if len(sys.argv) > 1 else 500
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 500
nrep = int(sys.argv[3]) if len(sys.argv) > 3 else 8
spec, body = configs.viv_cylinder_2d()
f0 = configs.uniform_state(spec, noise=1e-3)
f_host = f0.cpu().pin_memory()
b = dict(body, history=hist) if hist else dict(body)
ens = Ensemble([Stepper(spec, body=dict(b), dyn_mode="host") for _ in range(nrep)])
for st in ens.steppers:
    st.set_f(f_host)
ens.step(40)
for st in ens.steppers:
    st.get_f()
torch.cuda.synchronize()
for st in ens.steppers:
    st.set_f(f_host)
for k in range(4):
    t = time.perf_counter()
    try:
        ens.step(chunk)
    except Exception as exc:
        print("chunk", k, "FAILED", exc, "after", time.perf_counter() - t, "s")
        for st in ens.steppers:
            print(" mail", st._mail.tolist(), "pinned step", st._body_pin.numpy().view("int32")[22])
        raise
    if hist:
        rec = [st.body_history(min(hist, chunk) - 1) for st in ens.steppers]
    torch.cuda.synchronize()
    print("chunk", k, "ok", time.perf_counter() - t, "s", flush=True)
