#!/bin/bash
# Final visit of the round: parity tests, smoke, bench (+ reference arm).
set -u
TAG=${1:-r01final}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee $OUT/pytest_$TAG.log
echo "== smoke" ; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke_$TAG.log
echo "== bench" ; timeout 400 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err ; tail -c 1500 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
echo "== bench reference arm" ; timeout 200 python bench.py --impl reference --steps 200 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err ; cut -c1-600 $OUT/bench_ref_$TAG.json
