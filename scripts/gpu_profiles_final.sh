#!/bin/bash
# Profiles of the final round-1 code on one B200: parity tests, C3 with / without the IB chain, ncu launch lists of the
# bench command, C3 and C5, ncu --set full of the fused kernel (C2 bench) and of the tiled MDF stage (C5 body).
set -u
TAG=${1:-r01p}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee $OUT/pytest_$TAG.log
echo "== c3 chain on / off" ; timeout 200 python scripts/c3_nochain.py 2>&1 | tail -6 | tee $OUT/c3_nochain_$TAG.txt
echo "== ncu launch list of the bench command"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 -s 300 --csv --log-file $OUT/launches_bench_$TAG.csv \
    python bench.py --steps 36 --warmup 18 --no-extra --no-cpu-baseline > $OUT/ncu_bench_$TAG.log 2>&1
python scripts/launch_summary.py $OUT/launches_bench_$TAG.csv | head -6
echo "== ncu full capture of the fused kernel (C2 bench)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_step -s 60 -c 4 -f -o $OUT/prof_step2d_$TAG \
    python bench.py --steps 36 --warmup 18 --no-extra --no-cpu-baseline >> $OUT/ncu_bench_$TAG.log 2>&1
echo "== ncu launch list of C3"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 -s 40 --csv --log-file $OUT/launches_c3_$TAG.csv \
    python scripts/profile_kernels.py c3 6 > $OUT/ncu_c3_$TAG.log 2>&1
python scripts/launch_summary.py $OUT/launches_c3_$TAG.csv | head -4
echo "== ncu launch list of C5 on one GPU"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 -s 60 --csv --log-file $OUT/launches_c5_$TAG.csv \
    python scripts/config_runs.py --config c5 --steps 6 > $OUT/ncu_c5_$TAG.log 2>&1
python scripts/launch_summary.py $OUT/launches_c5_$TAG.csv | head -6
echo "== ncu full capture of the tiled MDF stage (C5 body)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_mdf_stage_tiled -s 9 -c 3 -f -o $OUT/prof_mdf_tiled_$TAG \
    python scripts/config_runs.py --config c5 --steps 3 >> $OUT/ncu_c5_$TAG.log 2>&1
ls -la $OUT/*.ncu-rep
