#!/bin/bash
# decoder attention on tcgen05: parity, stand-alone timing, c5x bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vqa.py -m gpu -q -x > gpurun_out/r2_vqa_tests.log 2>&1
echo "vqa tests rc=$?"; tail -n 25 gpurun_out/r2_vqa_tests.log
timeout 120 python tools/att_git_bench.py 2>&1 | tail -3
timeout 900 python bench.py --workload c5x --no-cpu-baseline > gpurun_out/r2_c5x_gitatt.json 2> gpurun_out/r2_c5x_gitatt.err
echo "c5x rc=$?"; python - <<PY
import json
d = json.load(open("gpurun_out/r2_c5x_gitatt.json"))
print(d["value"], d["ms_per_step"], d["roofline"].get("vqa_forward_ms_per_step"), (d.get("e2e") or {}).get("value"), d["clocks"])
PY
tail -n 3 gpurun_out/r2_c5x_gitatt.err
