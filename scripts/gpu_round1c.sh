#!/bin/bash
# GPU-box visit: host-mailbox e2e path (tests + bench) and the C3 breakdown.
set -u
TAG=${1:-r01c}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee $OUT/pytest_$TAG.log
echo "== bench" ; timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err ; python - <<PY
import json
d = json.load(open("$OUT/bench_$TAG.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "chunked", d["e2e"].get("chunked_device_ode", {}).get("value"), "roofline", d["roofline"]["frac"])
PY
tail -5 $OUT/bench_$TAG.err
echo "== c3 breakdown" ; timeout 600 python scripts/c3_breakdown.py 2>&1 | tee $OUT/c3_breakdown_$TAG.log
