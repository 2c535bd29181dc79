#!/bin/bash
# what the driver runs at round end, in its order: GPU tests, smoke, reference arm, bench (1 GPU)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 ) > gpurun_out/r2_final_tests.log 2>&1
echo "tests rc=$?"; tail -n 16 gpurun_out/r2_final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err
echo "ref rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_final_ref.json')); print(d['value'], d['cpu_baseline']['kind'], d['steps'])"
( time timeout 600 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err )
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_final_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['indices_identical'], d['status_counts'], d['clocks'])"
timeout 600 python bench.py --workload c2r --no-cpu-baseline > gpurun_out/r2_final_c2r.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_final_c2r.json')); print('c2r', d['value'], d['e2e']['value'])"
