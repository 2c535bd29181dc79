#!/bin/bash
# round 2, first GPU pass: the GPU test suite, the default bench line, the reference arm (1 GPU)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/r2_smi.log 2>&1
nproc > gpurun_out/r2_nproc.log
( time timeout 1200 python -m pytest tests -m gpu -q -x --durations=15 ) > gpurun_out/r2_tests.log 2>&1
echo "tests rc=$?"; tail -n 30 gpurun_out/r2_tests.log
timeout 600 python bench.py > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/r2_bench_c2.json; tail -n 5 gpurun_out/r2_bench_c2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
echo "ref rc=$?"; tail -c 1500 gpurun_out/r2_bench_ref.json; tail -n 3 gpurun_out/r2_bench_ref.err
