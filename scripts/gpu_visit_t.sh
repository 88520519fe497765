set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/pytest_r02t.log
echo "== single-domain parts"; VSB_PARTS_ONLY= timeout 600 python scripts/single_domain_parts.py barrier 2>&1 | grep "us per step" | tee $OUT/single_domain_parts_r02t.log
