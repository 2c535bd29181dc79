#!/bin/bash
# round 2 config matrix at N > 1 (run with gpurun --gpus N -- 'bash tools/gpu_r2_multi.sh N'):
#   c4 (BASELINE configs[3]: T=512, K=32) STRONG scaling over a fixed global list of 512 clips, at every N
#   c5x (configs[4]: exactly 10 000 clips -> K0 -> MDF -> GIT video-QA forward) and the NCCL sharding test at N = 8
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { # name port args...
  local name=$1 port=$2; shift 2
  timeout 900 $TR --master-port $port bench.py --gpus $N "$@" > gpurun_out/r2_$name.json 2> gpurun_out/r2_$name.err
  echo "$name rc=$?"; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_$name.json"))
    print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "scaling", "status_counts", "sharding_check")}, "e2e", (d.get("e2e") or {}).get("value"))
except Exception as e:
    print("no json:", e)
PY
  tail -n 3 gpurun_out/r2_$name.err
}
run c4_n$N 29521 --workload c4 --scaling strong --clips 512 --no-cpu-baseline
if [ "$N" = "8" ]; then
  run c5x_10k_n8 29522 --workload c5x --scaling strong --clips 10000 --no-cpu-baseline
  run c2_n8 29523 --workload c2 --no-cpu-baseline
  timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2_multi_test_n8.log 2>&1
  echo "multi test rc=$?"; tail -n 5 gpurun_out/r2_multi_test_n8.log
fi
