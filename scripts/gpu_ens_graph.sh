#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -12
for G in 1 0; do
  VSB_HOST_ODE_GRAPH=$G timeout 250 python bench.py --no-cpu-baseline --no-extra > $OUT/bench_g$G.json 2> $OUT/bench_g$G.err; tail -2 $OUT/bench_g$G.err
  python - <<PY
import json
d=json.loads(open("$OUT/bench_g$G.json").read().strip().splitlines()[-1])
print("graph=$G value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "single", round(d["e2e"]["single_domain_host_ode"]["value"]), "chunked", round(d["e2e"]["chunked_device_ode"]["value"]))
PY
done
VSB_HOST_ODE_THREADS=1 timeout 250 python bench.py --no-cpu-baseline --no-extra > $OUT/bench_g1t1.json 2> $OUT/bench_g1t1.err
python -c "
import json
d=json.loads(open('$OUT/bench_g1t1.json').read().strip().splitlines()[-1])
print('graph=1 threads=1 e2e', round(d['e2e']['value']))"
