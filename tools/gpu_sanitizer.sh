#!/bin/bash
# compute-sanitizer over a subset of the GPU suite (ONE tool per gpurun call: TOOL=memcheck | racecheck)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOL=${TOOL:-memcheck}
SEL='test_scan_env_schedule_bit_exact_vs_golden or test_feed_forward_generator_vs_golden or test_recurrent_generator_vs_golden or test_gather_atari_rows_bit_exact_vs_oracle or test_gather_empty_and_errors or test_fwd_ or test_dgrad_ or test_wgrad_ or test_s2d_and_colsum'
if [ "$TOOL" = memcheck ]; then SEL="$SEL or (test_gru_persistent_vs_torch and tc and (5-3-128 or 7-77-256))"; fi
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 97 --print-limit 20 \
  python -m pytest tests/test_scan_gather_gpu.py tests/test_tcx_gpu.py tests/test_engine_gpu.py -q -x -p no:cacheprovider -k "$SEL" \
  > gpurun_out/r02_sanitizer_$TOOL.log 2>&1
echo "sanitizer($TOOL) rc=$?"
grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|Error|hazard" gpurun_out/r02_sanitizer_$TOOL.log | tail -12
