#!/bin/bash
# First GPU visit of the next round (one B200, ~4 min of box time): everything that was added after the last GPU visit
# of round 1 gets its first hardware run here (tests/test_gpu_z_analytic.py: remove its xfail mark once the eight
# cases show up as XPASS), then the standard numbers.
#   gpurun --timeout 900 -- 'bash scripts/gpu_round2_first.sh r02a'
set -u
TAG=${1:-r02a}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu (no -x: see every failure)" ; timeout 400 python -m pytest tests -m gpu -q -rxX 2>&1 | tail -15 | tee $OUT/pytest_$TAG.log
echo "== smoke" ; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke_$TAG.log
echo "== bench" ; timeout 400 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err ; tail -c 1200 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
echo "== c3 chain on / off" ; timeout 200 python scripts/c3_nochain.py 2>&1 | tail -6 | tee $OUT/c3_nochain_$TAG.txt
echo "== c5 one GPU" ; timeout 200 python scripts/config_runs.py --config c5 --steps 6 2>&1 | tail -1 | cut -c1-330 | tee $OUT/c5_$TAG.log
echo "== fused kernel sweep" ; timeout 400 python scripts/vec_sweep.py 2>&1 | tee $OUT/vec_sweep_$TAG.txt
