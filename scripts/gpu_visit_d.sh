set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee $OUT/pytest_r02d.log
echo "== sweep moment"; VSB_MRT_FORM=moment timeout 300 python scripts/vec_sweep.py quick3d 2>&1 | tail -8 | tee $OUT/sweep_moment_r02d.log
echo "== sweep split"; VSB_MRT_FORM=split timeout 300 python scripts/vec_sweep.py quick3d 2>&1 | tail -8 | tee $OUT/sweep_split_r02d.log
echo "== chain probe"; timeout 600 python scripts/chain_probe.py 2>&1 | tail -12 | tee $OUT/chain_probe_r02d.log
echo "== bench"; timeout 900 python bench.py > $OUT/bench_r02d.json 2> $OUT/bench_r02d.err; tail -c 2500 $OUT/bench_r02d.json; tail -3 $OUT/bench_r02d.err
