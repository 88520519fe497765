set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== parts, 8 domains"; VSB_PARTS_DOMAINS=8 timeout 600 python scripts/single_domain_parts.py barrier launches cluster 2>&1 | grep "chain_first 0" | tee $OUT/ensemble_parts_r02f.log
