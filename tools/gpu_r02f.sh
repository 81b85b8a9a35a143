#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_engine_gpu.py -q -k "gru_persistent and tc-" > gpurun_out/r02f_gru_tc.log 2>&1
rc=$?; echo "gru tc (tensor-memory weights) tests rc=$rc" | tee -a gpurun_out/r02f_gru_tc.log
grep -E "passed|failed|Error|error|assert " gpurun_out/r02f_gru_tc.log | tail -12
A2CB_GRU_TC_DEBUG=1 timeout 300 python tools/gru_bench.py 128 64 512 > gpurun_out/r02f_gru_bench.jsonl 2>&1; echo "gru_bench rc=$?"; grep -E "tc|stamps" gpurun_out/r02f_gru_bench.jsonl | tail -4
if [ $rc -ne 0 ]; then
  export A2CB_GRU_TC_TMEM=0; echo "falling back to shared-memory weights for the rest of the session"
  A2CB_GRU_TC_DEBUG=1 timeout 300 python tools/gru_bench.py 128 64 512 > gpurun_out/r02f_gru_bench_smem.jsonl 2>&1; grep -E "tc|stamps" gpurun_out/r02f_gru_bench_smem.jsonl | tail -4
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02f_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_bench_c5.json 2> gpurun_out/r02f_bench_c5.err; echo "bench rc=$?"
python - <<'PY'
import json
r=json.load(open("gpurun_out/r02f_bench_c5.json"))
print("value %.0f e2e %.0f ms %.2f" % (r["value"], r["e2e"]["value"], r["ms_per_step"]))
for k in r["kernels"][:6]: print(k)
PY
