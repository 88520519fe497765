set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $OUT/pytest_r02h.log
echo "== c5 parts"; timeout 600 python scripts/c5_parts.py 2>&1 | tail -14 | tee $OUT/c5_parts_r02h2.log
echo "== bulk variants"; timeout 300 python scripts/bulk_variants.py 2>&1 | tail -8 | tee $OUT/bulk_variants_r02h.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_r02h.json 2> $OUT/bench_r02h.err; tail -c 600 $OUT/bench_r02h.json; tail -3 $OUT/bench_r02h.err
