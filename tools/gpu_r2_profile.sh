#!/bin/bash
# round 2 evidence (1 GPU): energy per kernel, ncu launch list of one step, full captures of the GEMM modes / attention /
# decoder attention / HBM kernels.  Each ncu capture only after the same command exited 0 without ncu.
mkdir -p gpurun_out
timeout 600 python tools/energy_bench.py --out gpurun_out/energy.json > gpurun_out/energy.log 2>&1
echo "energy rc=$?"; tail -n 4 gpurun_out/energy.log
SMALL="--clips 32 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
python bench.py $SMALL > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err || { echo "bench failed"; tail -5 gpurun_out/bench_small.err; exit 1; }
# one whole step = 2 chunks x (2 + 12 * 7) + pool/scores/select/gather launches; skip the 3 warm-up steps
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:sasvqa -s 546 -c 200 --csv \
    --log-file gpurun_out/launches.csv python bench.py $SMALL > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 60 -c 4 -f -o gpurun_out/prof_gemm \
    python bench.py $SMALL > gpurun_out/ncu_full.log 2>&1
echo "ncu gemm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 10 -c 1 -f -o gpurun_out/prof_att \
    python bench.py $SMALL > gpurun_out/ncu_att.log 2>&1
echo "ncu attention rc=$?"
ncu --set full --clock-control none -k regex:"layernorm_bf16|pool_norm|preprocess_u8|gather_u8|mdf_scores|pre_layernorm" -s 4 -c 8 -f -o gpurun_out/prof_hbm \
    python bench.py $SMALL > gpurun_out/ncu_hbm.log 2>&1
echo "ncu hbm rc=$?"
python tools/att_git_bench.py > gpurun_out/att_git_bench.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attention_git_tcgen05 -s 3 -c 1 -f -o gpurun_out/prof_att_git \
    python tools/att_git_bench.py > gpurun_out/ncu_att_git.log 2>&1
echo "ncu att_git rc=$?"; cat gpurun_out/att_git_bench.log
ls -la gpurun_out/*.ncu-rep
