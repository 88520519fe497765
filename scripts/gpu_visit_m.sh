set -u
OUT=gpurun_out; mkdir -p $OUT
for cfg in "2304 4 product" "2304 2 product" "1536 4 tile1536_c4" "1536 2 tile1536_c4" "1536 3 tile1536_c4"; do
  set -- $cfg
  if [ "$3" = "product" ]; then unset VIVSIM_B200_LIB; else export VIVSIM_B200_LIB=$PWD/build_variants/$3.so; fi
  VSB_TILE_CELLS=$1 VSB_TILE_COLUMN=$2 timeout 300 python scripts/tiled_probe.py 2>&1 | grep "ms per step" | tee -a $OUT/tiled_probe_r02m.log
done
