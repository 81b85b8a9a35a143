#!/bin/bash
# ncu --set full of the GRU tc kernels (one forward, one backward launch of tools/gru_bench.py)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/gru_bench.py 128 64 512 tc > gpurun_out/r02q_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gru_tc -s 6 -c 2 -f -o gpurun_out/r02q_gru python tools/gru_bench.py 128 64 512 tc > gpurun_out/r02q_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02q_ncu.log
