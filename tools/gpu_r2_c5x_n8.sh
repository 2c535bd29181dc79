#!/bin/bash
# BASELINE configs[4] on 8 GPUs with the final build: exactly 10 000 clips -> K0 -> MDF -> GIT video-QA forward
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 \
  --workload c5x --scaling strong --clips 10000 --no-cpu-baseline > gpurun_out/r2_c5x_10k_n8_final.json 2> gpurun_out/r2_c5x_10k_n8_final.err
echo "rc=$?"; python - <<PY
import json
d = json.load(open("gpurun_out/r2_c5x_10k_n8_final.json"))
print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "scaling", "sharding_check")}, "e2e", d["e2e"]["value"], d["per_rank"]["sm_mhz"], d["roofline"]["vqa_forward_ms_per_step"])
PY
tail -n 2 gpurun_out/r2_c5x_10k_n8_final.err
