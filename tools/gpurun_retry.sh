#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> <command...>   -- retries while the pod has no free GPU slot
to=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$to" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then
    echo "[retry $i] no slot yet"; sleep 90; continue
  fi
  echo "$out"; exit 0
done
echo "gave up"; exit 3
