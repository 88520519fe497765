#!/bin/bash
# Weak-scaling bench at N GPUs of one node:  gpurun --gpus N -- 'bash scripts/scale_check.sh N'
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  python bench.py --steps 20000 --warmup 200 --no-extra --no-cpu-baseline > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 20000 --warmup 200 2> gpurun_out/scale_$N.err | tail -1 > gpurun_out/scale_$N.json
fi
grep -E "Error|error|Traceback" gpurun_out/scale_$N.err | tail -5
python - <<PY
import json
d = json.load(open("gpurun_out/scale_$N.json"))
print("N=$N value", round(d["value"]), "MLUPS  ms/step", round(d["ms_per_step"], 5), " e2e", round(d["e2e"]["value"]),
      " per-GPU frac of HBM roofline", round(d["config"]["hbm_frac_of_measured"], 3))
print((d["config"].get("multi_gpu") or {}).get("exchange", "")[:80])
PY
