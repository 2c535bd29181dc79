"""Times the HBM-bound stage kernels alone and prints achieved GB/s (development aid, run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sasvqa_b200 import ops

PEAK = 6542.1


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
frames = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device="cuda")
ms = timeit(lambda: ops.preprocess_u8(frames))
print(f"preprocess_u8   {ms:.3f} ms  {n * 451584 / ms / 1e6:7.0f} GB/s  {n * 451584 / ms / 1e6 / PEAK:.2f} of peak")
B, T, K = n // 128, 128, 16
clips = frames.view(B, T, 224, 224, 3)
idx = torch.stack([torch.randperm(T, device="cuda")[:K] for _ in range(B)]).int()
ms = timeit(lambda: ops.gather_frames_u8(clips, idx))
by = B * K * (150528 + 602112)
print(f"gather_u8       {ms:.3f} ms  {by / ms / 1e6:7.0f} GB/s  {by / ms / 1e6 / PEAK:.2f} of peak")
for h, w in [(360, 640), (240, 320), (720, 1280)]:
    m = max(8, n // 8 if h < 700 else n // 32)
    raw = torch.randint(0, 256, (m, h, w, 3), dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: ops.resize_crop_u8(raw))
    full = m * (h * w * 3 + 150528)
    print(f"resize {h}x{w:<5d} {ms:.3f} ms  {m / ms * 1e3:9.0f} frames/s  {full / ms / 1e6:7.0f} GB/s of whole-frame bytes ({full / ms / 1e6 / PEAK:.2f})")
    del raw
x = torch.randn(n * 197, 768, device="cuda")
g, b = torch.randn(768, device="cuda"), torch.randn(768, device="cuda")
ms = timeit(lambda: ops.test_layernorm(x, g, b))
print(f"layernorm_bf16  {ms:.3f} ms  {n * 197 * 4608 / ms / 1e6:7.0f} GB/s  {n * 197 * 4608 / ms / 1e6 / PEAK:.2f} of peak")
feats = torch.nn.functional.normalize(torch.randn(B, T, 768, device="cuda"), dim=-1)
ms = timeit(lambda: ops.mdf_scores(feats, 8))
print(f"mdf_scores      {ms:.4f} ms  {B * T * 3076 / ms / 1e6:7.0f} GB/s")
del feats
# HBM-sized: 2048 clips x 128 frames of features = 805 MB (> L2), plus a long-video shape
for Bs, Ts, Ws in ((2048, 128, 8), (256, 512, 8), (2048, 128, 4)):
    big = torch.nn.functional.normalize(torch.randn(Bs, Ts, 768, device="cuda"), dim=-1)
    ms = timeit(lambda: ops.mdf_scores(big, Ws))
    gbs = Bs * Ts * 3076 / ms / 1e6
    print(f"mdf_scores B={Bs} T={Ts} W={Ws}  {ms:.4f} ms  {gbs:7.0f} GB/s  {gbs / PEAK:.2f} of peak")
    qb = torch.randn(Bs, 768, device="cuda")
    ms = timeit(lambda: ops.mif_scores(big, qb))
    gbs = Bs * Ts * 3076 / ms / 1e6
    print(f"mif_scores B={Bs} T={Ts}      {ms:.4f} ms  {gbs:7.0f} GB/s  {gbs / PEAK:.2f} of peak")
    del big
feats = torch.nn.functional.normalize(torch.randn(B, T, 768, device="cuda"), dim=-1)
q = torch.randn(B, 768, device="cuda")
ms = timeit(lambda: ops.mif_scores(feats, q))
print(f"mif_scores      {ms:.4f} ms  {B * T * 3076 / ms / 1e6:7.0f} GB/s")
