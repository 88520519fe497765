set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_r02k.log
echo "== c5 parts"; timeout 600 python scripts/c5_parts.py 2>&1 | tail -14 | tee $OUT/c5_parts_r02k.log
