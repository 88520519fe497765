set -u
OUT=gpurun_out; mkdir -p $OUT
for s in 5 6 1 2 3 4 0; do echo "== VSB_CLUSTER_STOP=$s"; VSB_CLUSTER_STOP=$s VSB_PARTS_ONLY=chain timeout 300 python scripts/single_domain_parts.py cluster 2>&1 | grep "chain_first 0" | tail -3; done | tee $OUT/cluster_stop_r02e.log
