#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
for T in 4 2 1; do
  VSB_HOST_ODE_THREADS=$T timeout 250 python bench.py --no-cpu-baseline --no-extra > $OUT/bench_thr$T.json 2> $OUT/bench_thr$T.err; tail -2 $OUT/bench_thr$T.err
  python - <<PY
import json
d=json.loads(open("$OUT/bench_thr$T.json").read().strip().splitlines()[-1])
print("threads=$T value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "single", round(d["e2e"]["single_domain_host_ode"]["value"]), "chunked", round(d["e2e"]["chunked_device_ode"]["value"]))
PY
done
