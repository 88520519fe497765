set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $OUT/pytest_r02g.log
echo "== bulk variants"; timeout 300 python scripts/bulk_variants.py 2>&1 | tail -8 | tee $OUT/bulk_variants_r02g.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_r02g.json 2> $OUT/bench_r02g.err; tail -c 1500 $OUT/bench_r02g.json; tail -3 $OUT/bench_r02g.err
echo "== ncu full c5 (moment MRT)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 4 -c 4 -o $OUT/ncu_full_c5_r02g -f python scripts/profile_kernels.py c5 3 > $OUT/ncu_full_c5_r02g.log 2>&1; tail -2 $OUT/ncu_full_c5_r02g.log
echo "== launch list c5"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_c5_r02g.csv python scripts/profile_kernels.py c5 3 > /dev/null 2>&1; tail -3 $OUT/launches_c5_r02g.csv | cut -c1-250
