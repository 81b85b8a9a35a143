#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r02x}
timeout 600 python bench.py --workload C4 --stacked-frames --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_C4.json 2> gpurun_out/${T}_bench_C4.err; echo "bench rc=$?"
python - <<PY
import json
r=json.load(open("gpurun_out/${T}_bench_C4.json"))
print("C4 value %.0f e2e %.0f ms %.3f launches %d" % (r["value"], r["e2e"]["value"], r["ms_per_step"], r["gpu_launches"]))
PY
A2CB_PROFILE_STEP=1 timeout 600 python bench.py --workload C4 --stacked-frames --steps 2 --warmup 12 --no-cpu-baseline > gpurun_out/${T}_plain.log 2>&1 &&
A2CB_PROFILE_STEP=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/${T}_launches_C4.csv python bench.py --workload C4 --stacked-frames --steps 2 --warmup 12 --no-cpu-baseline > gpurun_out/${T}_ncu.log 2>&1
echo "ncu rc=$?"
