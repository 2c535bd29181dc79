#!/bin/bash
# bench + ncu launch list + one full capture of the dominant kernel (run through gpurun, 1 GPU)
mkdir -p gpurun_out
python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err
rc=$?; echo "bench rc=$rc"; tail -c 3000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
[ $rc -eq 0 ] || exit $rc
# launch list: one whole step of the same command (skip the 3 warm-up steps)
SMALL="--clips 32 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
python bench.py $SMALL > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:sasvqa -s 546 -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py $SMALL > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
python bench.py $SMALL > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 60 -c 4 -o gpurun_out/prof_gemm \
    python bench.py $SMALL > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
