#!/bin/bash
# usage: tools/gpurun_retry_n.sh <gpus> <timeout> '<command>'  -- retries while the pod answers busy, up to ~45 min
N=$1; T=$2; shift; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --gpus "$N" --timeout "$T" -- "$@" > gpurun_out/.retry_n.log 2>&1
  rc=$?
  if ! grep -q "status=transient" gpurun_out/.retry_n.log; then tail -30 gpurun_out/.retry_n.log; exit $rc; fi
  sleep 60
done
echo "gave up: pod busy"; exit 3
