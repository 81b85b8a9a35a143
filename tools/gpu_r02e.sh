#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
A2CB_GRU_TC_DEBUG=1 timeout 300 python tools/gru_bench.py 128 64 512 > gpurun_out/r02e_gru_bench.jsonl 2>&1; echo "gru_bench rc=$?"; tail -8 gpurun_out/r02e_gru_bench.jsonl
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02e_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r02e_tests.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r02e_tests.log | tail -30
