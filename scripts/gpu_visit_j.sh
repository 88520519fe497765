set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_r02j.log
echo "== halo probe"; timeout 600 python scripts/halo_probe.py 2>&1 | tail -6 | tee $OUT/halo_probe_r02j.log
