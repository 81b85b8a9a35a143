#!/bin/bash
# round-2 GPU session B: new GRU kernels first (under a timeout: cooperative kernels that wait on each other), then the suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_engine_gpu.py -q -k "gru_persistent and tc" > gpurun_out/r02b_gru_tc.log 2>&1
rc=$?; echo "gru tc tests rc=$rc" | tee -a gpurun_out/r02b_gru_tc.log
tail -15 gpurun_out/r02b_gru_tc.log
if [ $rc -ne 0 ]; then
  A2CB_GRU_TC_SWAP=1 timeout 300 python -m pytest tests/test_engine_gpu.py -q -k "gru_persistent and tc" > gpurun_out/r02b_gru_tc_swap.log 2>&1
  echo "gru tc tests (swapped MN strides) rc=$?" | tee -a gpurun_out/r02b_gru_tc_swap.log
  tail -15 gpurun_out/r02b_gru_tc_swap.log
fi
timeout 300 python tools/gru_bench.py 128 64 512 > gpurun_out/r02b_gru_bench.jsonl 2>&1; echo "gru_bench rc=$?"; cat gpurun_out/r02b_gru_bench.jsonl
if [ $rc -ne 0 ]; then export A2CB_GRU_MODE=1; echo "falling back to the mma.sync GRU for the rest of the session"; fi
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r02b_tests.log
tail -8 gpurun_out/r02b_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench_c5.json 2> gpurun_out/r02b_bench_c5.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02b_bench_c5.err
for w in C1 C2 C3 C4; do timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r02b_bench_$w.json 2> gpurun_out/r02b_bench_$w.err; echo "$w rc=$?"; done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
