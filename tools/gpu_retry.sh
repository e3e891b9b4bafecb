#!/bin/bash
# usage: tools/gpu_retry.sh <timeout-seconds> [--gpus N] <command> : gpurun with retries while the pod's GPU slots are busy (exit 3)
T=$1; shift
OPTS=()
if [ "$1" = "--gpus" ]; then OPTS=(--gpus "$2"); shift 2; fi
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun --timeout "$T" "${OPTS[@]}" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
