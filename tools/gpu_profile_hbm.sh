mkdir -p gpurun_out
SMALL="--clips 32 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
python bench.py $SMALL > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err || { echo "bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:sasvqa -s 546 -c 200 --csv \
    --log-file gpurun_out/launches_v8.csv python bench.py $SMALL > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
ncu --set full --clock-control none -k regex:"pool_norm|preprocess_u8|gather_u8|mdf_scores|mdf_greedy|topk" -s 4 -c 8 -f -o gpurun_out/prof_hbm_v8 \
    python bench.py $SMALL > gpurun_out/ncu_hbm.log 2>&1
echo "ncu hbm rc=$?"
