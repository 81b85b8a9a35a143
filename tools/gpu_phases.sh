#!/bin/bash
# where does a sharded step spend its time?  N = 1 and N = 2 on the SAME box, phase marks (A2CB_PHASE_TIMING=1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r02zb}
A2CB_PHASE_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-parity > gpurun_out/${T}_phases_n1.json 2> gpurun_out/${T}_phases_n1.err; echo "n1 rc=$?"
grep "^phases" gpurun_out/${T}_phases_n1.err | sed -n 4p | tr "|" "\n" | head -30
A2CB_PHASE_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/${T}_phases_n2.json 2> gpurun_out/${T}_phases_n2.err; echo "n2 rc=$?"
echo "--- balanced (resident)"; grep "^phases" gpurun_out/${T}_phases_n2.err | sed -n 4p | tr "|" "\n" | head -40
echo "--- balanced (e2e)"; grep "^phases" gpurun_out/${T}_phases_n2.err | sed -n 8p | tr "|" "\n" | head -40
echo "--- stratified"; grep "^phases" gpurun_out/${T}_phases_n2.err | tail -1 | tr "|" "\n" | head -40
python - <<PY
import json
for n in (1, 2):
    r=json.loads(open("gpurun_out/${T}_phases_n%d.json" % n).read().strip().splitlines()[-1])
    print("n%d value %.0f e2e %.0f ms %.3f" % (n, r["value"], r["e2e"]["value"], r["ms_per_step"]), r.get("stratified_sharding",{}).get("ms_per_step"))
PY
