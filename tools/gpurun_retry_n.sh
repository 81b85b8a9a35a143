#!/bin/bash
# usage: tools/gpurun_retry_n.sh <gpus> <timeout-seconds> <command...>   -- retries while the pod has no free slot of that size
g=$1; to=$2; shift; shift
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --gpus "$g" --timeout "$to" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then
    echo "[retry $i] no slot yet"; sleep 60; continue
  fi
  echo "$out"; exit 0
done
echo "gave up"; exit 3
