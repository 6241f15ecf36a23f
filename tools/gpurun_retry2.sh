#!/bin/bash
# usage: tools/gpurun_retry2.sh <gpus> <timeout> '<command>'
G=$1; T=$2; shift; shift
for i in $(seq 1 14); do
  /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@" > gpurun_out/.retry.log 2>&1
  rc=$?
  if ! grep -q "status=transient" gpurun_out/.retry.log; then cat gpurun_out/.retry.log | tail -25; exit $rc; fi
  sleep 45
done
echo "gave up: pod busy"; exit 3
