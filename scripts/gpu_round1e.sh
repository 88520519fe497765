#!/bin/bash
# GPU-box visit: cell-centric tiled MDF stages.  Tests, C5 on one GPU (tiled / untiled) and its launch list.
set -u
TAG=${1:-r01e}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee $OUT/pytest_$TAG.log
echo "== c5 one GPU: tiled / untiled"
timeout 600 python scripts/config_runs.py --config c5 --steps 6 2>&1 | tail -1 | tee $OUT/c5_$TAG.log
VSB_MDF_UNTILED=1 timeout 600 python scripts/config_runs.py --config c5 --steps 6 2>&1 | tail -1 | tee -a $OUT/c5_$TAG.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 -s 60 --csv --log-file $OUT/launches_c5_$TAG.csv \
    python scripts/config_runs.py --config c5 --steps 3 > $OUT/ncu_c5_$TAG.log 2>&1
python scripts/launch_summary.py $OUT/launches_c5_$TAG.csv 2>&1 | head -6
