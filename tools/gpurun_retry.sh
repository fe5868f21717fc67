#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> '<command>' : retries while the pod answers "transient" (no slot free)
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "gave up: pod busy"; exit 3
