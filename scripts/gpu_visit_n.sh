set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee $OUT/pytest_r02n.log
echo "== single-domain parts"; timeout 600 python scripts/single_domain_parts.py cta barrier 2>&1 | grep "us per step" | tee $OUT/single_domain_parts_r02n.log
