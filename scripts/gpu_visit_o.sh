set -u
OUT=gpurun_out; mkdir -p $OUT
for s in 1 2 3 4 5 6 0; do VSB_CTA_STOP=$s timeout 120 python scripts/cta_chain_phases.py cta 2>&1 | grep "us per step" | tee -a $OUT/cta_phases_r02o.log; done
timeout 120 python scripts/cta_chain_phases.py barrier 2>&1 | grep "us per step" | tee -a $OUT/cta_phases_r02o.log
