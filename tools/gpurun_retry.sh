#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <script> [gpus] - run a GPU call, retrying while the pod has no slot free (nothing is charged for those)
for i in $(seq 1 20); do
  if [ -n "$3" ]; then out=$(/usr/local/graft/bin/gpurun --gpus "$3" --timeout "$1" -- bash "$2" 2>&1); else out=$(/usr/local/graft/bin/gpurun --timeout "$1" -- bash "$2" 2>&1); fi
  echo "$out" | tail -45
  echo "$out" | grep -q "status=transient" || exit 0
  sleep 120
done
