#!/bin/bash
# per-kernel durations of one L2-resident C2 domain for a chain mode:  bash scripts/ncu_chain.sh cluster
MODE=${1:-cluster}
OUT=gpurun_out; mkdir -p $OUT
cat > /tmp/one_domain.py <<PY
import sys, torch
sys.path.insert(0, ".")
from vivsim_b200 import Stepper, configs
spec, body = configs.viv_cylinder_2d()
st = Stepper(spec, body=dict(body), dyn_mode="device", ib_chain="$MODE")
st.set_f(configs.uniform_state(spec, noise=1e-3)); st.step(41); torch.cuda.synchronize()
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file $OUT/launches_chain_$MODE.csv python /tmp/one_domain.py > $OUT/launches_chain_$MODE.log 2>&1
python scripts/launch_summary.py $OUT/launches_chain_$MODE.csv 2>&1 | tail -12
