#!/bin/bash
# two-GPU session: real-NCCL parity check, then the C5 and C3 benches at N = 2 (weak scaling)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > gpurun_out/r02z_dist_check.log 2>&1; echo "dist_check rc=$?"; tail -4 gpurun_out/r02z_dist_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02z_bench_C5_n2.json 2> gpurun_out/r02z_bench_C5_n2.err; echo "C5 n2 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload C3 --steps 20 --warmup 3 > gpurun_out/r02z_bench_C3_n2.json 2> gpurun_out/r02z_bench_C3_n2.err; echo "C3 n2 rc=$?"
python - <<'PY'
import json
for f in ("C5_n2","C3_n2"):
    try:
        r=json.loads(open("gpurun_out/r02z_bench_%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.0f e2e %.0f ms %.3f" % (r["value"], r["e2e"]["value"], r["ms_per_step"]), r["config"].get("recurrent_sharding"))
    except Exception as e:
        print(f, "failed", e)
PY
