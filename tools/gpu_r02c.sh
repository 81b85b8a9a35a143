#!/bin/bash
# round-2 GPU session C: tcgen05 GRU kernels (under a timeout), whole suite, C5 bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_engine_gpu.py -q -k "gru_persistent and tc-" > gpurun_out/r02c_gru_tc.log 2>&1
rc=$?; echo "gru tc tests rc=$rc" | tee -a gpurun_out/r02c_gru_tc.log
grep -E "passed|failed|Error|error" gpurun_out/r02c_gru_tc.log | tail -12
if [ $rc -ne 0 ]; then
  A2CB_GRU_TC_SWAP=1 timeout 300 python -m pytest tests/test_engine_gpu.py -q -k "gru_persistent and tc-" > gpurun_out/r02c_gru_tc_swap.log 2>&1
  echo "gru tc tests (swapped MN strides) rc=$?" | tee -a gpurun_out/r02c_gru_tc_swap.log
  grep -E "passed|failed|Error|error|assert" gpurun_out/r02c_gru_tc_swap.log | tail -12
fi
timeout 300 python tools/gru_bench.py 128 64 512 > gpurun_out/r02c_gru_bench.jsonl 2>&1; echo "gru_bench rc=$?"; cat gpurun_out/r02c_gru_bench.jsonl | tail -5
if [ $rc -ne 0 ]; then export A2CB_GRU_MODE=1; echo "falling back to the mma.sync GRU for the rest of the session"; fi
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02c_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r02c_tests.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r02c_tests.log | tail -30
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02c_bench_c5.json 2> gpurun_out/r02c_bench_c5.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02c_bench_c5.err
timeout 600 python bench.py --workload C4 --steps 10 --warmup 3 --stacked-frames --no-cpu-baseline > gpurun_out/r02c_bench_C4_stacked.json 2> gpurun_out/r02c_bench_C4_stacked.err; echo "C4 stacked rc=$?"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
