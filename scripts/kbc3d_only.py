import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vivsim_b200 import Stepper, configs
spec = dict(dim=3, shape=(256, 256, 256), collision=sys.argv[1] if len(sys.argv) > 1 else "kbc", omega=1.7, forcing=None, post=[], u0=0.05)
st = Stepper(spec).set_f(configs.uniform_state(spec, noise=1e-3)); st.step(4); torch.cuda.synchronize()
