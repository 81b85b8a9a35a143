#!/bin/bash
# N-GPU session (GPUS=4 | 8): C5 bench with per-step times
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
G=${GPUS:-8}
A2CB_BENCH_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --steps 10 --warmup 3 > gpurun_out/r02z_bench_C5_n$G.json 2> gpurun_out/r02z_bench_C5_n$G.err; echo "rc=$?"
grep -E "^step" gpurun_out/r02z_bench_C5_n$G.err | sed "s/step \([0-9]*\) (\([a-z0-9]*\)): /\2:/" | tr "\n" " "; echo
python - <<PY
import json
r=json.loads(open("gpurun_out/r02z_bench_C5_n$G.json").read().strip().splitlines()[-1])
print("C5 n$G value %.0f e2e %.0f ms %.3f" % (r["value"], r["e2e"]["value"], r["ms_per_step"]), r["config"].get("recurrent_sharding"), r.get("stratified_sharding",{}).get("value"))
PY
tail -5 gpurun_out/r02z_bench_C5_n$G.err | cut -c1-300
