#!/bin/bash
# where does a sharded step spend its time?  N = 1 and N = 2 on the SAME box, phase marks (A2CB_PHASE_TIMING=1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r02zb}
G=${GPUS:-2}
for c in ${CASES:-C5s C5m}; do
A2CB_DIST_STAGED=1 A2CB_DIST_CASE=$c timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${CHECK_GPUS:-2} --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > gpurun_out/${T}_dist_check_$c.log 2>&1; echo "dist_check $c rc=$?"; tail -1 gpurun_out/${T}_dist_check_$c.log | cut -c1-200
done
if [ "${SKIP_N1:-0}" != "1" ]; then
A2CB_PHASE_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-parity > gpurun_out/${T}_phases_n1.json 2> gpurun_out/${T}_phases_n1.err; echo "n1 rc=$?"
echo "--- n1 resident"; grep "^phases" gpurun_out/${T}_phases_n1.err | sed -n 4p | tr "|" "\n" | head -7
echo "--- n1 e2e"; grep "^phases" gpurun_out/${T}_phases_n1.err | sed -n 8p | tr "|" "\n" | head -7
fi
A2CB_PHASE_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --steps 3 --warmup 2 > gpurun_out/${T}_phases_n$G.json 2> gpurun_out/${T}_phases_n$G.err; echo "n$G rc=$?"
echo "--- balanced (resident)"; grep "^phases" gpurun_out/${T}_phases_n$G.err | sed -n 4p | tr "|" "\n" | head -12
echo "--- balanced (e2e)"; grep "^phases" gpurun_out/${T}_phases_n$G.err | sed -n 8p | tr "|" "\n" | head -12
echo "--- stratified"; grep "^phases" gpurun_out/${T}_phases_n$G.err | tail -1 | tr "|" "\n" | head -4
python - <<PY
import json, os
for n in (1, $G):
    f = "gpurun_out/${T}_phases_n%d.json" % n
    if not os.path.exists(f): continue
    r=json.loads(open(f).read().strip().splitlines()[-1])
    print("n%d value %.0f e2e %.0f ms %.3f" % (n, r["value"], r["e2e"]["value"], r["ms_per_step"]), r.get("stratified_sharding",{}).get("ms_per_step"))
PY
