#!/bin/bash
# usage: tools/gpu.sh <timeout-seconds> '<command>'   -- retries while the pod's GPU slots are busy (exit code 3)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
