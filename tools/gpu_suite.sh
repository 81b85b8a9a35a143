#!/bin/bash
# GPU session: tcx / engine / update tests, then the C5 bench (TAG names the outputs)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r02x}
timeout 900 python -m pytest tests/test_tcx_gpu.py tests/test_engine_gpu.py tests/test_update_gpu.py -q -x > gpurun_out/${T}_tests.log 2>&1; rc=$?
echo "tests rc=$rc"; tail -4 gpurun_out/${T}_tests.log
if [ $rc -ne 0 ]; then grep -E "^E |Error|FAILED" gpurun_out/${T}_tests.log | head -20; exit 1; fi
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_c5.json 2> gpurun_out/${T}_bench_c5.err; echo "bench rc=$?"
python - <<PY
import json
r=json.load(open("gpurun_out/${T}_bench_c5.json"))
print("value %.0f e2e %.0f ms %.2f parity %s" % (r["value"], r["e2e"]["value"], r["ms_per_step"], r.get("parity",{}).get("max_rel_err")))
for k in r["kernels"][:16]: print("%-14s %.3f ms %.3f %s" % (k["op"], k["ms"], k["share"], k["tflops"]))
PY
