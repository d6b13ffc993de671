#!/bin/bash
# like gpurun_retry.sh with --gpus N: tools/gpurun_retry_n.sh <N> <timeout_s> '<command>'
cd "$(dirname "$0")/.."
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus "$1" --timeout "$2" -- "$3"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
