#!/bin/bash
# ncu captures of the MIF cross-encoder (bench.py --workload c3x) for profiles/: launch list of one step + full
# captures of its kernels (run through gpurun on ONE GPU; each capture only after the same command exited 0 without ncu)
mkdir -p gpurun_out
SMALL="--workload c3x --clips 64 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
python bench.py $SMALL > gpurun_out/bench_c3x_small.json 2> gpurun_out/bench_c3x_small.err || { echo "bench failed"; tail -5 gpurun_out/bench_c3x_small.err; exit 1; }
# 3 warm-up steps x 100 launches (1 embed + 12 x 8 + pooler + label/topk) come first; list the 4th (timed) step
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:sasvqa -s 300 -c 100 --csv \
    --log-file gpurun_out/launches_c3x.csv python bench.py $SMALL > gpurun_out/ncu_list_c3x.log 2>&1
echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"attention_short|layernorm_post|embed_layernorm|pooler_classifier" -s 13 -c 6 -f -o gpurun_out/prof_scorer_small \
    python bench.py $SMALL > gpurun_out/ncu_scorer_small.log 2>&1
echo "ncu scorer kernels rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 48 -c 4 -f -o gpurun_out/prof_scorer_gemm \
    python bench.py $SMALL > gpurun_out/ncu_scorer_gemm.log 2>&1
echo "ncu scorer gemm rc=$?"
ls -la gpurun_out/*.ncu-rep
