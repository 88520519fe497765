#!/bin/bash
# One GPU visit (one B200): tests, smoke, the driver's bench command, the reference arm, launch list.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_visit.sh r02a'
set -u
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu (no -x: see every failure)" ; timeout 900 python -m pytest tests -m gpu -q -rxXs 2>&1 | tail -40 | tee $OUT/pytest_$TAG.log
echo "== smoke" ; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke_$TAG.log
if [ -n "${EXTRA:-}" ]; then echo "== extra: $EXTRA"; timeout 600 bash -c "$EXTRA" 2>&1 | tail -60 | tee $OUT/extra_$TAG.log; fi
echo "== bench (driver command)" ; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err ; tail -c 3000 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
echo "== reference arm" ; timeout 300 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > $OUT/bench_ref_$TAG.json 2>&1 ; tail -c 600 $OUT/bench_ref_$TAG.json
