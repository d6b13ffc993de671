#!/bin/bash
# Retry gpurun while the pod answers "busy" (exit code 3: nothing charged).  usage: tools/gpurun_retry.sh <timeout_s> '<command>'
cd "$(dirname "$0")/.."
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
