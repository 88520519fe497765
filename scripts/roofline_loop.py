"""The dominant kernel of the C2 step alone, rotating over 8 domains (working set 604 MB > L2), launched eagerly so that
ncu sees every launch in steady state:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none \
        --clock-control none -k regex:k_step -s 64 -c 32 --csv python scripts/roofline_loop.py

(one pass per launch for these three metrics, so the caches keep their steady-state content: the DRAM bytes include the
write-back of the previous launches' output, which a cold-cache capture of a single launch misses)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vivsim_b200 import Stepper, configs  # noqa: E402

spec, _ = configs.viv_cylinder_2d()
plain = dict(spec); plain.pop("ib"); plain["post"] = []
f0 = configs.uniform_state(spec, noise=1e-3)
ks = [Stepper(plain).set_f(f0) for _ in range(8)]
for s in ks:
    s.step(1)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 16):
    for s in ks:
        s.advance_raw(1)
torch.cuda.synchronize()
print("done")
