#!/bin/bash
# ncu --set full of the MDF chain kernel of one C2 domain:  bash scripts/ncu_full_chain.sh cluster k_mdf_cluster2d
MODE=${1:-cluster}; KERN=${2:-k_mdf_cluster2d}
OUT=gpurun_out; mkdir -p $OUT
cat > /tmp/one_domain.py <<PY
import sys, torch
sys.path.insert(0, ".")
from vivsim_b200 import Stepper, configs
spec, body = configs.viv_cylinder_2d()
st = Stepper(spec, body=dict(body), dyn_mode="device", ib_chain="$MODE")
st.set_f(configs.uniform_state(spec, noise=1e-3)); st.step(21); torch.cuda.synchronize()
PY
timeout 300 ncu --set full --import-source on --clock-control none -k regex:$KERN -s 6 -c 1 -f -o $OUT/ncu_full_chain_$MODE python /tmp/one_domain.py > $OUT/ncu_full_chain_$MODE.log 2>&1
tail -3 $OUT/ncu_full_chain_$MODE.log
