#!/bin/bash
# Quick check of a kernel change on one B200: parity tests, C3 chain on/off, C5 on one GPU, C2 bench summary.
set -u
TAG=${1:-q}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee $OUT/pytest_$TAG.log
echo "== c3" ; timeout 200 python scripts/c3_nochain.py 2>&1 | tail -6 | tee $OUT/c3_nochain_$TAG.txt
echo "== c5" ; timeout 200 python scripts/config_runs.py --config c5 --steps 6 2>&1 | tail -1 | cut -c1-330 | tee $OUT/c5_$TAG.log
echo "== bench" ; timeout 300 python bench.py --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
print("C2 value", round(d["value"]), "frac", round(d["config"]["hbm_frac_of_measured"],3), "l2res", round(d["config"]["l2_resident_mlups"]), "e2e", round(d["e2e"]["value"]), "kernel frac", round(d["roofline"]["frac"],3), [ (a["workload"][:2], round(a["mlups"]), round(a["hbm_frac_of_measured"],3)) for a in d.get("also",[])])
PY
