#!/bin/bash
# gpurun with retries while the pod answers "transient" (exit 3 / nothing charged).  Usage: tools/gpurun_retry.sh <timeout> '<command>'
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out"
  exit 0
done
echo "$out"
exit 3
