#!/bin/bash
# round-2 multi-GPU session: real-NCCL checks of the sharded updates (fused exchange) + the bench at N ranks
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
for c in C3 C5s C5m C2 C4s; do
  A2CB_DIST_CASE=$c run tests/dist_check.py > gpurun_out/r02d_dist_${c}_n$N.log 2>&1; echo "dist_check $c n=$N rc=$?"; tail -2 gpurun_out/r02d_dist_${c}_n$N.log
done
run bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02d_bench_c5_n$N.json 2> gpurun_out/r02d_bench_c5_n$N.err; echo "bench n=$N rc=$?"
cat gpurun_out/r02d_bench_c5_n$N.json | cut -c 1-700
run bench.py --gpus $N --steps 10 --warmup 3 --recurrent-sharding stratified > gpurun_out/r02d_bench_c5_n${N}_strat.json 2> gpurun_out/r02d_bench_c5_n${N}_strat.err; echo "bench stratified n=$N rc=$?"
cat gpurun_out/r02d_bench_c5_n${N}_strat.json | cut -c 1-400
