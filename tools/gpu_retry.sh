#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> '<command>'   -- retries while the pod has no free GPU slot (gpurun exit 3)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" "${GPUS_ARG[@]}" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpu_retry] no slot (try $i), sleeping 60 s"
  sleep 60
done
exit 3
