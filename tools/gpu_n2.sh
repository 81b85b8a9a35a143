#!/bin/bash
# two-GPU session: real-NCCL parity checks (C3 feed-forward, C5s / C5m recurrent with the default balanced sharding),
# then the C5 bench at N = 2 (weak scaling)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in C3 C5s C5m; do
  A2CB_DIST_CASE=$c timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > gpurun_out/r02z_dist_check_$c.log 2>&1; echo "dist_check $c rc=$?"; tail -1 gpurun_out/r02z_dist_check_$c.log
done
A2CB_DIST_STAGED=1 A2CB_DIST_CASE=C5m timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > gpurun_out/r02z_dist_check_C5m_staged.log 2>&1; echo "dist_check C5m staged rc=$?"; tail -1 gpurun_out/r02z_dist_check_C5m_staged.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02z_bench_C5_n2.json 2> gpurun_out/r02z_bench_C5_n2.err; echo "C5 n2 rc=$?"
python - <<'PY'
import json
r=json.loads(open("gpurun_out/r02z_bench_C5_n2.json").read().strip().splitlines()[-1])
print("C5 n2 value %.0f e2e %.0f ms %.3f" % (r["value"], r["e2e"]["value"], r["ms_per_step"]), r["config"].get("recurrent_sharding"), r.get("stratified_sharding",{}).get("value"))
PY
