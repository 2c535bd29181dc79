#!/bin/bash
# round 2, second GPU pass (1 GPU): full GPU suite, then the N=1 legs of the config matrix with e2e everywhere
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=12 ) > gpurun_out/r2_tests2.log 2>&1
echo "tests rc=$?"; tail -n 40 gpurun_out/r2_tests2.log
run() { # name args...
  local name=$1; shift
  timeout 900 python bench.py "$@" > gpurun_out/r2_$name.json 2> gpurun_out/r2_$name.err
  echo "$name rc=$?"; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_$name.json"))
    print({k: d.get(k) for k in ("value", "ms_per_step", "status_counts")}, "e2e", (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("clips_per_gpu_per_step"), d["roofline"].get("energy"))
except Exception as e:
    print("no json:", e)
PY
  tail -n 3 gpurun_out/r2_$name.err
}
run c4_n1 --workload c4 --scaling strong --clips 512 --no-cpu-baseline
run c5x_n1 --workload c5x --scaling strong --clips 1250 --no-cpu-baseline
run c3_n1 --workload c3 --no-cpu-baseline
run c2r_n1 --workload c2r --no-cpu-baseline
run c5_n1 --workload c5 --clips 320 --no-cpu-baseline
