#!/bin/bash
# end-of-session check on one GPU: the whole -m gpu suite, smoke(), then the default bench line (CPU baseline included)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r02zf}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${T}_bench_C5.json 2> gpurun_out/${T}_bench_C5.err; echo "bench rc=$?"
python - <<PY
import json
r=json.load(open("gpurun_out/${T}_bench_C5.json"))
print("value %.0f e2e %.0f ms %.3f parity %s cpu %s" % (r["value"], r["e2e"]["value"], r["ms_per_step"], r.get("parity",{}).get("max_rel_err"), r.get("cpu_baseline",{}).get("value")))
print(r["roofline"]["op"], r["roofline"]["frac"], r["clocks"])
PY
