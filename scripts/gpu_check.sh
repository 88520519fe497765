#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list + full capture of the fused kernel.
# Usage (from the repo root): gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [tag]'
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu_$TAG.txt 2>&1
echo "== pytest -m gpu" ; python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $OUT/pytest_$TAG.log
echo "== smoke" ; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke_$TAG.log
echo "== bench" ; python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err ; tail -c 3000 $OUT/bench_$TAG.json ; tail -5 $OUT/bench_$TAG.err
echo "== bench reference arm" ; python bench.py --impl reference --steps 200 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err ; cat $OUT/bench_ref_$TAG.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 -s 300 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 36 --warmup 18 --no-extra --no-cpu-baseline > $OUT/ncu_bench_$TAG.log 2>&1
echo "== ncu full capture of the fused kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 60 -c 4 -f -o $OUT/prof_step2d_$TAG \
    python bench.py --steps 36 --warmup 18 --no-extra --no-cpu-baseline >> $OUT/ncu_bench_$TAG.log 2>&1
ls -la $OUT | tail -20
