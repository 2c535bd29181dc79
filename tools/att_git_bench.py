"""Times the GIT decoder attention kernel alone (test hook) -- development aid, run on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sasvqa_b200 import _capi

n, n_vis, L = 16, 16 * 197, 20
S = n_vis + L
qkv = torch.randn(n * S, 2304, device="cuda").to(torch.bfloat16)
out = torch.empty(n * S, 768, dtype=torch.bfloat16, device="cuda")
fn = lambda: _capi.check(_capi.lib().sasvqa_test_attention_git(qkv.data_ptr(), n, n_vis, L, out.data_ptr(),
                                                               torch.cuda.current_stream().cuda_stream), "attention_git")
for _ in range(3):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
flop = n * 12 * 4.0 * (n_vis * n_vis + L * n_vis + L * (L + 1) / 2) * 64
print(f"attention_git: {ms:.3f} ms for {n} samples of {n_vis}+{L} rows  ({flop / ms / 1e9:.0f} TFLOP/s useful)")
