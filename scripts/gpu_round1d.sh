#!/bin/bash
# GPU-box visit: inactive-warp early exit in k_step + tiled MDF stages.  Tests, bench, C3 breakdown, C5 on one GPU
# (tiled vs untiled) and its launch list.
set -u
TAG=${1:-r01d}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee $OUT/pytest_$TAG.log
echo "== bench" ; timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err ; python - <<PY
import json
d = json.load(open("$OUT/bench_$TAG.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "chunked", d["e2e"].get("chunked_device_ode", {}).get("value"), "roofline", d["roofline"]["frac"])
for a in d.get("also", []): print(a["workload"][:40], a["mlups"], a["hbm_frac_of_measured"])
PY
tail -5 $OUT/bench_$TAG.err
echo "== c3 breakdown" ; timeout 600 python scripts/c3_breakdown.py 2>&1 | tee $OUT/c3_breakdown_$TAG.log
echo "== c5 one GPU: tiled / untiled"
timeout 600 python scripts/config_runs.py --config c5 --steps 6 2>&1 | tail -1 | tee $OUT/c5_$TAG.log
VSB_MDF_UNTILED=1 timeout 600 python scripts/config_runs.py --config c5 --steps 6 2>&1 | tail -1 | tee -a $OUT/c5_$TAG.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 -s 30 --csv --log-file $OUT/launches_c5_$TAG.csv \
    python scripts/config_runs.py --config c5 --steps 3 > $OUT/ncu_c5_$TAG.log 2>&1
python scripts/launch_summary.py $OUT/launches_c5_$TAG.csv 2>&1 | tail -8
