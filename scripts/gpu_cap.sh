#!/bin/bash
# MDF grid-cap sweep on C3 (+ parity tests under a cap).
set -u
OUT=gpurun_out; mkdir -p $OUT
for CAP in 0 148 296 444; do
  echo "== VSB_MDF_GRID_CAP=$CAP"
  VSB_MDF_GRID_CAP=$CAP timeout 200 python scripts/c3_breakdown.py 2>&1 | grep -E "body only|full C3  |full C3, overlap off" | tee -a $OUT/c3_cap_sweep.log
done
echo "== tests with cap 148"; VSB_MDF_GRID_CAP=148 timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "== tests with cap 7"; VSB_MDF_GRID_CAP=7 timeout 300 python -m pytest tests/test_gpu_step.py -m gpu -q -x 2>&1 | tail -4
