#!/bin/bash
# L2-residency probe: the same 64-clip step at different chunk sizes -- ms, joules, per-stage ms
for c in 31 62 124 248 2048; do
  timeout 600 python bench.py --clips 64 --chunk-frames $c --no-e2e --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r2_chunk_$c.json 2> gpurun_out/r2_chunk_$c.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r2_chunk_$c.json"))
r = d["roofline"]
st = r["stage_ms_per_step"]
g = sum(v for k, v in st.items() if k.startswith("gemm_"))
print("chunk $c: %.1f ms/step  %.1f videos/s  %s J  gemm %.1f ms  att %.1f  ln %.1f  launches %d  clk %s" % (d["ms_per_step"], d["value"], r["energy"]["joules_per_step"], g, st["attention"], st["layernorm"], d["gpu_launches"], d["clocks"]["sm_mhz"]))
PY
done
