#!/bin/bash
# Runs on the B200 box (through gpurun).  Each stage is its own process with its own timeout so a
# trap in one kernel cannot take the other results down with it.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
run() { # name timeout cmd...
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.log
  timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc" | tee -a gpurun_out/summary.log
  tail -n 25 "gpurun_out/$name.log" | tee -a gpurun_out/summary.log
}
: > gpurun_out/summary.log
run t1_nogemm 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "simt or layernorm or attention or preprocess or gather or scores or select or topk" --maxfail=20
run t2_tcgen05 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tcgen05" --maxfail=20
run t4_pipeline 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "encoder or e2e or edge or full_size" --maxfail=20
run t5_smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
