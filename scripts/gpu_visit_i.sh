set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== overlap probe"; timeout 900 python scripts/overlap_probe.py 2>&1 | tail -10 | tee $OUT/overlap_probe_r02i.log
