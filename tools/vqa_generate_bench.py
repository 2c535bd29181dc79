"""Times greedy answer decoding (sas.vqa_generate) -- development aid, run on the GPU box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sasvqa_b200 as sas
from sasvqa_b200 import synth

B, K, L0, Lmax = 64, 16, 20, 50
enc = sas.FrameEncoder(synth.random_encoder_state_dict(synth.REF_SEED), chunk_frames=2048)
psd = synth.random_projection_state_dict()
enc.set_projection(*[psd[f"visual_projection.{k}"] for k in ("0.weight", "0.bias", "1.weight", "1.bias")])
dec = sas.GitDecoder(synth.random_git_decoder_state_dict(), max_rows=131072)
frames = torch.randn(B, K, 3, 224, 224, device="cuda")
ids = torch.randint(1000, synth.GIT_VOCAB, (B, L0), device="cuda")
for fn, name in ((lambda: sas.vqa_logits(frames, ids, enc, dec), "vqa_logits (one forward, 20 text rows)"),
                 (lambda: sas.vqa_generate(frames, ids, enc, dec, max_length=Lmax, trim=False), f"vqa_generate ({Lmax - L0} greedy steps)")):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"{name}: {dt * 1e3:.1f} ms for {B} samples of {K} frames = {B / dt:.0f} samples/s")
