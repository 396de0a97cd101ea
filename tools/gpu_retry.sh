#!/bin/bash
# usage: tools/gpu_retry.sh <timeout-s> '<command>'  -- retries while the pod answers "busy" (nothing charged)
t=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$t" -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out"; exit $rc
done
echo "gave up: pod busy"; exit 3
