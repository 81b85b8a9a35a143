#!/bin/bash
# two-GPU session: real-NCCL parity checks (C3 feed-forward, C5s / C5m recurrent with the default balanced sharding; gradient
# exchange through the peer-memory all-reduce, and through NCCL for comparison), then the C5 bench at N = 2 (weak scaling)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r02z}
run_check() {  # name, extra env
  env $2 A2CB_DIST_CASE=$1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > gpurun_out/${T}_dist_check_$1$3.log 2>&1; echo "dist_check $1 $2 rc=$?"; tail -1 gpurun_out/${T}_dist_check_$1$3.log | cut -c1-300
}
run_check C3 "A2CB_X=0" ""
run_check C5s "A2CB_X=0" ""
run_check C5m "A2CB_DIST_STAGED=1" "_staged"
run_check C5m "A2CB_EXCHANGE=nccl" "_nccl"
run_check C5m "A2CB_P2P_WAIT=spin" "_spin"
for mode in p2p nccl; do
A2CB_EXCHANGE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps ${STEPS:-8} --warmup 3 > gpurun_out/${T}_bench_C5_n2_$mode.json 2> gpurun_out/${T}_bench_C5_n2_$mode.err; echo "C5 n2 $mode rc=$?"
python - <<PY
import json
r=json.loads(open("gpurun_out/${T}_bench_C5_n2_$mode.json").read().strip().splitlines()[-1])
print("C5 n2 $mode value %.0f e2e %.0f ms %.3f" % (r["value"], r["e2e"]["value"], r["ms_per_step"]), r["config"].get("recurrent_sharding"), r["config"].get("exchange"), r.get("stratified_sharding",{}).get("value"))
PY
done
