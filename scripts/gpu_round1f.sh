#!/bin/bash
# Re-entry check of HEAD on one B200: parity tests, smoke, bench, C5 and C3 on one GPU.
# Usage: gpurun --timeout 840 -- 'bash scripts/gpu_round1f.sh [tag]'
set -u
TAG=${1:-r01f}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" ; ( time timeout 420 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) 2>&1 | tee $OUT/pytest_$TAG.log
echo "== smoke" ; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_$TAG.log
echo "== bench" ; ( time timeout 300 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err ) 2>&1 | tail -3 ; tail -c 2500 $OUT/bench_$TAG.json ; tail -3 $OUT/bench_$TAG.err
echo "== c5 one GPU" ; timeout 200 python scripts/config_runs.py --config c5 --steps 6 2>&1 | tail -2 | tee $OUT/c5_$TAG.log
echo "== c3 breakdown" ; timeout 200 python scripts/c3_breakdown.py 2>&1 | tee $OUT/c3_breakdown_$TAG.log
