#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>'   (retries while the pod's GPU slots are busy)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then
    sleep 90
    continue
  fi
  cat /tmp/gpurun_last.log
  exit $rc
done
cat /tmp/gpurun_last.log
exit 3
