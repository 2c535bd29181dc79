#!/bin/bash
# final verification of the merged build: whole GPU suite, smoke, default bench, c5x, launch list of one c5x step
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/r2_final2_tests.log 2>&1
echo "tests rc=$?"; tail -n 14 gpurun_out/r2_final2_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r2_final2_bench.json 2> gpurun_out/r2_final2_bench.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_final2_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['indices_identical'], d['clocks'])"
timeout 900 python bench.py --workload c5x --no-cpu-baseline > gpurun_out/r2_final2_c5x.json 2> gpurun_out/r2_final2_c5x.err
echo "c5x rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_final2_c5x.json')); print(d['value'], d['e2e']['value'], d['roofline']['vqa_forward_ms_per_step'], d['clocks'])"
SMALL="--workload c5x --clips 16 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
python bench.py $SMALL > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:sasvqa -c 4000 --csv \
    --log-file gpurun_out/launches_c5x.csv python bench.py $SMALL > gpurun_out/ncu_list_c5x.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/launches_c5x.csv
