#!/bin/bash
# usage: [NG=2] tools/gpu_retry.sh <timeout-seconds> '<command>'   -- retries while the pod answers "busy" (exit code 3)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun ${NG:+--gpus $NG} --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
