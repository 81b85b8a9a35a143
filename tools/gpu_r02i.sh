#!/bin/bash
# round-2 GPU session I: GRU tc kernels with merged hi/lo state operand (2 MMAs per reduction step), control warp
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r02i}
timeout 300 python -m pytest tests/test_engine_gpu.py -q -x -k "gru_persistent" > gpurun_out/${T}_gru_tc.log 2>&1
rc=$?; echo "gru tests rc=$rc" | tee -a gpurun_out/${T}_gru_tc.log
grep -E "passed|failed|Error|error|assert " gpurun_out/${T}_gru_tc.log | tail -12
A2CB_GRU_TC_DEBUG=1 timeout 300 python tools/gru_bench.py 128 64 512 > gpurun_out/${T}_gru_bench.jsonl 2>&1; echo "gru_bench rc=$?"; grep -E "tc|stamps|Error" gpurun_out/${T}_gru_bench.jsonl | tail -4
if [ $rc -ne 0 ]; then exit 1; fi
timeout 600 python -m pytest tests/test_update_gpu.py tests/test_engine_gpu.py -q -x -k "C5 or recurrent or gru" > gpurun_out/${T}_tests.log 2>&1; echo "recurrent tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_c5.json 2> gpurun_out/${T}_bench_c5.err; echo "bench rc=$?"
python - <<PY
import json
r=json.load(open("gpurun_out/${T}_bench_c5.json"))
print("value %.0f e2e %.0f ms %.2f parity %s" % (r["value"], r["e2e"]["value"], r["ms_per_step"], r.get("parity",{}).get("max_rel_err")))
for k in r["kernels"][:8]: print(k)
PY
