#!/bin/bash
# benches of every workload (C5 with the CPU baseline), then the launch list + ncu --set full of one timed C5 step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02y_bench_C5.json 2> gpurun_out/r02y_bench_C5.err; echo "C5 rc=$?"
for w in C1 C2 C3 C4; do
  extra=""; if [ $w = C4 ] || [ $w = C2 ]; then extra="--stacked-frames"; fi
  timeout 600 python bench.py --workload $w $extra --steps 20 --warmup 3 > gpurun_out/r02y_bench_$w.json 2> gpurun_out/r02y_bench_$w.err; echo "$w rc=$?"
done
python - <<'PY'
import json
for w in ("C5","C1","C2","C3","C4"):
    try:
        r=json.load(open("gpurun_out/r02y_bench_%s.json"%w))
        print(w, "value %.0f e2e %.0f ms %.3f cpu %.0f / %.0f (1 thread)" % (r["value"], r["e2e"]["value"], r["ms_per_step"], r["cpu_baseline"]["value"], r["cpu_baseline"].get("threads_1",{}).get("value",0)))
    except Exception as e:
        print(w, "failed", e)
PY
A2CB_PROFILE_STEP=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02y_plain.log 2>&1 &&
A2CB_PROFILE_STEP=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/r02y_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02y_ncu.log 2>&1
echo "ncu launches rc=$?"
A2CB_PROFILE_STEP=1 timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:gru_tc|tcx_fwd|tcx_wgrad' -c 19 -f \
   -o gpurun_out/r02y_c5_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02y_ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/r02y_c5_full.ncu-rep
