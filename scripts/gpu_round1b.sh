#!/bin/bash
# GPU-box visit of the second session of round 1: parity tests (incl. post / multigrid / history / checkpoint), smoke,
# bench, and ncu launch lists of the 3-D configurations (C3 sphere, C5 cylinder on one GPU).
# Usage: gpurun --timeout 1200 -- 'bash scripts/gpu_round1b.sh [tag]'
set -u
TAG=${1:-r01b}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" ; python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee $OUT/pytest_$TAG.log
echo "== smoke" ; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke_$TAG.log
echo "== bench" ; python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err ; tail -c 1500 $OUT/bench_$TAG.json ; tail -5 $OUT/bench_$TAG.err
echo "== c3 / c5 timed" ; python scripts/profile_kernels.py c3 20 2>&1 | tail -2 | tee $OUT/c3_$TAG.log
python scripts/config_runs.py --config c5 --steps 6 2>&1 | tail -2 | tee $OUT/c5_$TAG.log
echo "== ncu launch lists"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 -s 40 --csv --log-file $OUT/launches_c3_$TAG.csv \
    python scripts/profile_kernels.py c3 6 > $OUT/ncu_c3_$TAG.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 -s 30 --csv --log-file $OUT/launches_c5_$TAG.csv \
    python scripts/config_runs.py --config c5 --steps 3 > $OUT/ncu_c5_$TAG.log 2>&1
python scripts/launch_summary.py $OUT/launches_c3_$TAG.csv 2>&1 | tail -12
python scripts/launch_summary.py $OUT/launches_c5_$TAG.csv 2>&1 | tail -12
ls -la $OUT | tail -20
