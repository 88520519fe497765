#!/bin/bash
# Multi-GPU visit:  gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_multi.sh N tag'
N=${1:-2}; TAG=${2:-r02}
OUT=gpurun_out; mkdir -p $OUT
if [ "$N" = "2" ]; then echo "== pytest -m gpu (2 GPUs: includes tests/test_gpu_multi.py)"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $OUT/pytest_${N}gpu_$TAG.log; fi
echo "== multi_gpu_worker on $N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$([ $N -gt 4 ] && echo 4 || echo $N) --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_worker.py > $OUT/multi_$TAG.log 2>&1; grep "^\[\|MULTI" $OUT/multi_$TAG.log | tail -24
echo "== bench --gpus $N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_${N}gpu_$TAG.json 2> $OUT/bench_${N}gpu_$TAG.err; tail -c 2500 $OUT/bench_${N}gpu_$TAG.json; grep -v -i warn $OUT/bench_${N}gpu_$TAG.err | tail -8
