#!/bin/bash
# A/B of decoder-attention variants (stand-alone timing + parity of the attention test on each)
for s in "" _p4 _p2 ""; do
  echo "== variant '$s'"
  SASVQA_LIB_SUFFIX=$s timeout 120 python tools/att_git_bench.py 2>&1 | tail -1
  SASVQA_LIB_SUFFIX=$s timeout 300 python -m pytest tests/test_gpu_vqa.py -m gpu -q -k git_attention 2>&1 | tail -1
done
