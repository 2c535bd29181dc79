"""Per-call latency of the caption cross-encoder at the reference's own batch size (one QA sample = the K captions of
one video, gen_sample.py:79-88) -- development aid, run on the GPU box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sasvqa_b200 as sas
from sasvqa_b200 import synth

scorer = sas.CaptionScorer(synth.random_scorer_state_dict(vocab=4096), max_tokens=8192)
tok = synth.SynthTokenizer(4096)
for T in (16, 32, 128):
    qa, caps = synth.make_qa_workload(1, T, seed=1)
    b = tok(text=[qa[0]["question"]] * T, text_pair=caps["video0"])
    for _ in range(5):
        scorer.select_captions_host(b["input_ids"], b["token_type_ids"], b["attention_mask"], 1, 8, 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 50
    for _ in range(reps):
        idx, _ = scorer.select_captions_host(b["input_ids"], b["token_type_ids"], b["attention_mask"], 1, 8, 1)
    dt = (time.perf_counter() - t0) / reps
    print(f"{T} captions ({int(b['attention_mask'].sum())} tokens): {dt * 1e3:.3f} ms per QA sample through "
          f"sasvqa_mif_select_captions_host (ids in, 8 indices out) = {1 / dt:.0f} samples/s one at a time")
scorer.close()
