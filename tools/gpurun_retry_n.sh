#!/bin/bash
# like gpurun_retry.sh with a GPU count: tools/gpurun_retry_n.sh <gpus> <timeout> '<command>'
G=$1; T=$2; shift 2
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
