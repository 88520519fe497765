set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== bench"; timeout 1200 python bench.py --steps 20 --warmup 5 > $OUT/bench_r02p1.json 2> $OUT/bench_r02p1.err; tail -c 1500 $OUT/bench_r02p1.json; grep -v -i warn $OUT/bench_r02p1.err | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
