"""Sharding invariance on real GPUs (pytest -m gpu, needs >= 2 devices; skipped on a one-GPU box): the index tables
gathered over NCCL from rank-sharded clip lists / QA lists are identical to a single-GPU run over the whole list
(SURVEY.md section 4 (iii)).  One process per GPU, rendezvous on 127.0.0.1."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["SASVQA_ROOT"])
import sasvqa_b200 as sas
from sasvqa_b200 import sharding, synth
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
enc = sas.FrameEncoder(synth.random_encoder_state_dict(synth.REF_SEED), chunk_frames=128)
n, T, K, W = 7, 40, 5, 3
make = lambda s, e: synth.make_clips(range(100 + s, 100 + e), T, device=dev)
res = sharding.sample_mdf_sharded(make, n, enc, K, W)
whole = sas.sample_mdf_batch(make(0, n), enc, K, W)
assert torch.equal(res["indices"], whole["indices"]), (rank, res["indices"], whole["indices"])
assert torch.equal(res["status"], whole["status"])
# a ragged clip list, sharded by frame count, against one ragged call over the whole list on this rank
lens = [30, 7, 52, 12, 19, 41]
rag = [synth.make_clip(200 + i, t) for i, t in enumerate(lens)]
rs = sharding.sample_mdf_ragged_sharded(rag, enc, K, W)
whole_r = sas.sample_mdf_ragged([c.to(dev) for c in rag], enc, K, W)
assert torch.equal(rs["indices"], whole_r["indices"]) and torch.equal(rs["status"], whole_r["status"]), (rank, rs["spans"])
# the MIF step, QA list sharded by rank, against the unsharded call on this rank
scorer = sas.CaptionScorer(synth.random_scorer_state_dict(vocab=2048), max_tokens=4096)
tok = synth.SynthTokenizer(2048)
qa, caps = synth.make_qa_workload(9, 12, seed=3)
sharded = sharding.generate_inds_sharded(tok, scorer, qa, caps, 4, 1, samples_per_call=2)
single = sas.generate_inds(tok, scorer, qa, caps, 4, 1)
assert [r["sampled_inds"] for r in sharded] == [r["sampled_inds"] for r in single]
enc.close(); scorer.close()
dist.barrier()
dist.destroy_process_group()
print("OK", rank)
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_tables_equal_single_gpu(tmp_path):
    world = min(torch.cuda.device_count(), 4)
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, SASVQA_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29671", WORLD_SIZE=str(world))
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"OK {r}" in o, o[-3000:]
