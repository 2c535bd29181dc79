"""Times the encoder GEMM shapes alone (test hook) -- development aid, run on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sasvqa_b200 import ops, _capi
M = int(sys.argv[1]) if len(sys.argv) > 1 else 2048 * 197
shapes = [("qkv", 2304, 768, 0), ("out", 768, 768, 2), ("fc1", 3072, 768, 1), ("fc2", 768, 3072, 2)]
tot = 0.0
line = []
for name, N, K, mode in shapes:
    a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    vec = torch.randn(N, device="cuda")
    x = torch.zeros(M, N, device="cuda") if mode == 2 else None
    for _ in range(3):
        ops.test_gemm(a, b, mode, vec, out_f32=x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = int(os.environ.get("REPS", "20"))
    e0.record()
    for _ in range(reps):
        ops.test_gemm(a, b, mode, vec, out_f32=x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tot += ms
    line.append(f"{name} {ms:.3f} ms {2 * M * N * K / ms / 1e9:.0f} TF/s")
    del a, b, x
print(os.path.basename(_capi.LIB_PATH), " | ".join(line), f"| sum {tot:.3f} ms")
