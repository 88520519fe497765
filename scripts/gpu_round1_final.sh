#!/bin/bash
# Final 1-GPU visit of round 1: parity tests, smoke, bench (+ reference arm), ncu launch list of the bench command,
# ncu --set full of the fused kernel, fused-kernel sweep, C3 breakdown, C5 on one GPU.
# Usage: gpurun --timeout 2400 -- 'bash scripts/gpu_round1_final.sh [tag]'
set -u
TAG=${1:-r01z}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu_$TAG.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $OUT/pytest_$TAG.log
echo "== smoke" ; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_$TAG.log
echo "== bench" ; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err ; tail -c 1200 $OUT/bench_$TAG.json ; tail -3 $OUT/bench_$TAG.err
echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 200 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err ; cat $OUT/bench_ref_$TAG.json
echo "== fused kernel sweep" ; timeout 600 python scripts/vec_sweep.py 2>&1 | tee $OUT/vec_sweep_$TAG.txt
echo "== c3 breakdown" ; timeout 600 python scripts/c3_breakdown.py 2>&1 | tee $OUT/c3_breakdown_$TAG.log
echo "== c5 one GPU" ; timeout 600 python scripts/config_runs.py --config c5 --steps 6 2>&1 | tail -1 | tee $OUT/c5_$TAG.log
echo "== ncu launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 -s 300 --csv --log-file $OUT/launches_bench_$TAG.csv \
    python bench.py --steps 36 --warmup 18 --no-extra --no-cpu-baseline > $OUT/ncu_bench_$TAG.log 2>&1
python scripts/launch_summary.py $OUT/launches_bench_$TAG.csv | head -8
echo "== ncu full capture of the fused kernel (C2 bench)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 60 -c 4 -f -o $OUT/prof_step2d_$TAG \
    python bench.py --steps 36 --warmup 18 --no-extra --no-cpu-baseline >> $OUT/ncu_bench_$TAG.log 2>&1
echo "== ncu launch list of C3"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 -s 40 --csv --log-file $OUT/launches_c3_$TAG.csv \
    python scripts/profile_kernels.py c3 6 > $OUT/ncu_c3_$TAG.log 2>&1
python scripts/launch_summary.py $OUT/launches_c3_$TAG.csv | head -4
ls -la $OUT | tail -12
