set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest (chain modes)"; timeout 600 python -m pytest tests/test_gpu_step.py -m gpu -x -q -k "chain or rotating" 2>&1 | tail -6 | tee $OUT/pytest_r02q.log
echo "== single-domain parts"; timeout 600 python scripts/single_domain_parts.py cluster_barrier barrier 2>&1 | grep "us per step" | tee $OUT/single_domain_parts_r02q.log
