#!/bin/bash
# gpurun with retries while the pod answers "transient" (no box / slot free; nothing charged)
# usage: scripts/gpu_retry.sh <timeout-seconds> '<command>'
T=$1; shift
for i in $(seq 1 40); do
  out=$(gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
