#!/bin/bash
# c4 (BASELINE configs[3]) strong scaling 1 -> 2 -> 4 -> 8 on ONE 8-GPU box (same silicon for every N)
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --workload c4 --scaling strong --clips 512 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r2_onebox_c4_n1.json 2> gpurun_out/r2_onebox_c4_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29530 + N)) bench.py --gpus $N --workload c4 --scaling strong --clips 512 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r2_onebox_c4_n$N.json 2> gpurun_out/r2_onebox_c4_n$N.err
  fi
  echo "N=$N rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/r2_onebox_c4_n$N.json"))
print(d["n_gpus"], round(d["value"], 2), round(d["e2e"]["value"], 2), d["clocks"]["sm_mhz"], (d.get("per_rank") or {}).get("sm_mhz"), (d.get("sharding_check") or {}).get("identical"))
PY
done
