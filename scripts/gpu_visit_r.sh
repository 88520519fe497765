set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== ncu --set full of the grid-barrier chain (C2)"
VSB_CHAIN=barrier timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mdf_stage -s 10 -c 2 -o $OUT/ncu_full_chain_r02r -f \
  python scripts/profile_kernels.py c2 8 > $OUT/ncu_full_chain_r02r.log 2>&1
tail -2 $OUT/ncu_full_chain_r02r.log
