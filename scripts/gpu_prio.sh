#!/bin/bash
# IB-stream priority experiment: C2 bench, C3 breakdown and C5 with the IB chain on a high-priority stream vs default.
set -u
TAG=${1:-r01g}
OUT=gpurun_out
mkdir -p $OUT
for P in -1 0; do
  export VSB_IB_PRIORITY=$P
  echo "== VSB_IB_PRIORITY=$P"
  timeout 200 python bench.py --no-cpu-baseline > $OUT/bench_prio${P}_$TAG.json 2> $OUT/bench_prio${P}_$TAG.err
  python - <<PY
import json
d=json.loads(open("$OUT/bench_prio${P}_$TAG.json").read().strip().splitlines()[-1])
print("C2 value", round(d["value"]), "frac", round(d["config"]["hbm_frac_of_measured"],3), "l2res", round(d["config"]["l2_resident_mlups"]), "e2e", round(d["e2e"]["value"]), [ (a["workload"][:2], round(a["mlups"]), round(a["hbm_frac_of_measured"],3)) for a in d.get("also",[])])
PY
  timeout 200 python scripts/c3_breakdown.py 2>&1 | grep -E "body only|full C3" | tee $OUT/c3_prio${P}_$TAG.log
  timeout 200 python scripts/config_runs.py --config c5 --steps 6 2>&1 | tail -1 | cut -c1-330 | tee $OUT/c5_prio${P}_$TAG.log
done
