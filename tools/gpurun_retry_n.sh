#!/bin/bash
# like gpurun_retry.sh with --gpus N: tools/gpurun_retry_n.sh <N> <timeout_s> '<command>'
N=$1; T=$2; shift; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --gpus "$N" --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
