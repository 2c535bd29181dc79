"""Development aid: one traced attention launch per variant (needs the -DSASVQA_ATT_TRACE build)."""
import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 2:
    for v in sys.argv[1:]:
        print(f"=== variant {v}", flush=True)
        subprocess.run([sys.executable, __file__, v])
    sys.exit(0)
import torch
from sasvqa_b200 import ops
v = int(sys.argv[1])
qkv = (torch.randn(2048 * 197, 2304, device="cuda") * 1.5).to(torch.bfloat16)
for _ in range(4):
    ops.test_attention(qkv, impl=10 + v)
torch.cuda.synchronize()
