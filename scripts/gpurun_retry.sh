#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3):  scripts/gpurun_retry.sh <timeout_s> '<command>' [gpus]
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 12); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$CMD"; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$CMD"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] busy, attempt $i; sleeping 90 s"; sleep 90
done
exit 3
