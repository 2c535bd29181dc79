#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_suite.py -m gpu -q -x -s -k full_size ) > gpurun_out/r2_fullc2.log 2>&1
echo "rc=$?"; grep -E "full-size c2|passed|failed|real" gpurun_out/r2_fullc2.log | tail -5
