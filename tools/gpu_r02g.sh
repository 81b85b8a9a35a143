#!/bin/bash
# round-2 GPU session G: whole GPU suite, C5 bench with the CPU baseline, launch list of one timed step (ncu, shares only)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r02g_tests.log
tail -5 gpurun_out/r02g_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02g_bench_c5.json 2> gpurun_out/r02g_bench_c5.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02g_bench_c5.err
A2CB_PROFILE_STEP=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02g_plain.log 2>&1 &&
A2CB_PROFILE_STEP=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/r02g_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02g_ncu.log 2>&1
echo "ncu rc=$?"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
