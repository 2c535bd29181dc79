#!/bin/bash
# probes: NVDEC library on the box, compute-sanitizer on small cases
ls /usr/lib/x86_64-linux-gnu 2>/dev/null | grep -i -E "nvcuvid|nvidia-encode|libcuda" | head
python - <<'PY'
import ctypes
for n in ("libnvcuvid.so.1", "libnvcuvid.so", "libnvidia-encode.so.1"):
    try:
        ctypes.CDLL(n); print(n, "loads")
    except OSError as e:
        print(n, "absent:", str(e)[:80])
PY
echo "NVIDIA_DRIVER_CAPABILITIES=$NVIDIA_DRIVER_CAPABILITIES"
which compute-sanitizer
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "select or scores or preprocess or gather or layernorm" 2>&1 | tail -8
echo "memcheck small rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
echo "memcheck smoke rc=$?"
