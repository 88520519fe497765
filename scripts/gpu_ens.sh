#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for O in 1 0; do
  VSB_BENCH_OVERLAP=$O timeout 200 python bench.py --no-cpu-baseline --no-extra > $OUT/bench_ens$O.json 2> $OUT/bench_ens$O.err
  python - <<PY
import json
d=json.loads(open("$OUT/bench_ens$O.json").read().strip().splitlines()[-1])
print("overlap=$O C2 value", round(d["value"]), "frac", round(d["config"]["hbm_frac_of_measured"],3), "us/step", round(d["ms_per_step"]*1e3,2), "launches", d["gpu_launches"])
PY
done
