"""Times the attention kernels alone (test hook) -- development aid, run on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sasvqa_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
impls = [int(a) for a in sys.argv[2:]] or [0, 1]
qkv = (torch.randn(n * 197, 2304, device="cuda") * 1.5).to(torch.bfloat16)
for impl in impls:
    for _ in range(3):
        ops.test_attention(qkv, impl=impl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        ops.test_attention(qkv, impl=impl)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    items = n * 12
    print(f"impl {impl}: {ms:.3f} ms for {n} frames  ({ms * 1e-3 / items * 148 * 1e9:.0f} ns per item per SM; "
          f"{4 * 197 * 197 * 64 * items / ms / 1e9:.0f} TFLOP/s useful)")
