#!/bin/bash
# ncu launch list of the full video-QA forward (bench.py --workload c5x, small): which kernels the decoder spends its time in
mkdir -p gpurun_out
SMALL="--workload c5x --clips 16 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
python bench.py $SMALL > gpurun_out/bench_c5x_small.json 2> gpurun_out/bench_c5x_small.err || { echo "bench failed"; tail -5 gpurun_out/bench_c5x_small.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:sasvqa --csv \
    --log-file gpurun_out/launches_c5x.csv python bench.py $SMALL > gpurun_out/ncu_list_c5x.log 2>&1
echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attention_git -s 8 -c 2 -f -o gpurun_out/prof_att_git \
    python bench.py $SMALL > gpurun_out/ncu_att_git.log 2>&1
echo "ncu attention_git rc=$?"
