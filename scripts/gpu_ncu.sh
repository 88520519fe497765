#!/bin/bash
# ncu captures of one visit (one B200): steady-state DRAM traffic of the dominant kernel, the launch list of the bench
# command, --set full of the shipped D3Q19 kernels.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_ncu.sh r02'
set -u
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
echo "== steady-state DRAM traffic of k_step<2,BGK,vec4>"
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none \
  -k regex:k_step -s 64 -c 32 --csv --log-file $OUT/traffic_$TAG.csv python scripts/roofline_loop.py > $OUT/traffic_$TAG.log 2>&1
tail -4 $OUT/traffic_$TAG.csv
echo "== launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_bench_$TAG.csv \
  python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline > $OUT/launches_bench_$TAG.log 2>&1
tail -3 $OUT/launches_bench_$TAG.csv | cut -c1-300
for cfg in c3 c5; do
  echo "== ncu --set full $cfg"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 4 -c 4 -o $OUT/ncu_full_${cfg}_$TAG -f \
    python scripts/profile_kernels.py $cfg 3 > $OUT/ncu_full_${cfg}_$TAG.log 2>&1
  tail -2 $OUT/ncu_full_${cfg}_$TAG.log
done
