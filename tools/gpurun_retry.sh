#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> '<command>'  -- retries while the pod answers busy (exit 3), up to ~40 min
T=$1; shift
for i in $(seq 1 14); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > gpurun_out/.retry.log 2>&1
  rc=$?
  if ! grep -q "status=transient" gpurun_out/.retry.log; then cat gpurun_out/.retry.log | tail -25; exit $rc; fi
  sleep 45
done
echo "gave up: pod busy"; exit 3
