set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== ncu --set full of k_mdf_stage_tiled (C5 recipe at 256^3)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mdf_stage_tiled -s 6 -c 3 -o $OUT/ncu_full_tiled_r02l -f \
  python scripts/profile_kernels.py c5 3 > $OUT/ncu_full_tiled_r02l.log 2>&1
tail -2 $OUT/ncu_full_tiled_r02l.log
