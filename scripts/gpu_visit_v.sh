set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_r02v.log
for pf in 0 1; do
  echo "== c3 parts, VSB_PREFETCH_BOX=$pf"; VSB_PREFETCH_BOX=$pf timeout 300 python scripts/c5_parts.py c3 2>&1 | grep "ms per step" | sed "s/^/box prefetch $pf: /" | tee -a $OUT/c3_parts_r02v.log
done
for pf in 0 1; do
  echo "== c5 whole step, VSB_PREFETCH_BOX=$pf"; VSB_PREFETCH_BOX=$pf timeout 300 python scripts/tiled_probe.py 2>&1 | grep "ms per step" | sed "s/^/box prefetch $pf: /" | tee -a $OUT/c3_parts_r02v.log
done
