#!/bin/bash
# ncu --set full capture of the tiled MDF stage on the C5 body (one GPU).
set -u
TAG=${1:-r01h}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python scripts/config_runs.py --config c5 --steps 6 2>&1 | tail -1 | tee $OUT/c5_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mdf_stage_tiled -s 9 -c 3 -f -o $OUT/prof_mdf_tiled_$TAG \
    python scripts/config_runs.py --config c5 --steps 3 > $OUT/ncu_tiled_$TAG.log 2>&1
ls -la $OUT/*.ncu-rep
