"""GPU parity tests of the downstream video-QA forward (pytest -m gpu): MyGitForCausalLM's inference logits
(src/modeling/modeling.py:29-232) through the C ABI against the fp32 oracle (oracle/git.py, pinned to HF's
GitForCausalLM by tests/golden/git_vqa_hf.npz) on the same seeded weights, frames and token ids.

Tolerances: decoder attention vs fp32 |d| <= 2e-2 on bf16 outputs; text embeddings (fp32) |d| <= 2e-5; final hidden state
cosine >= 0.999 per row; logits |d| <= 6e-2 on logits of std 1.1 (bf16 through 12 encoder + 6 decoder blocks and a
768-long head dot product), top-1 token identical unless the fixture's top-2 margin is below 2 * that row's error.
"""
import os

import numpy as np
import pytest
import torch

from oracle import git as git_oracle, vit
import sasvqa_b200 as sas
from sasvqa_b200 import _capi, ops, synth, vqa

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def weights():
    return (synth.random_encoder_state_dict(synth.REF_SEED), synth.random_projection_state_dict(),
            synth.random_git_decoder_state_dict())


@pytest.fixture(scope="module")
def models(weights):
    torch.cuda.set_device(0)
    enc_sd, psd, dsd = weights
    enc = ops.FrameEncoder(enc_sd, chunk_frames=64)
    enc.set_projection(*[psd[f"visual_projection.{k}"] for k in ("0.weight", "0.bias", "1.weight", "1.bias")])
    dec = vqa.GitDecoder(dsd, max_rows=4096)
    yield enc, dec
    dec.close()
    enc.close()


@pytest.mark.parametrize("n,n_vis,L,ramp", [(1, 197, 1, 0.0), (2, 394, 9, 0.0), (3, 130, 40, 0.0), (1, 64, 200, 0.0),
                                            (1, 256, 5, 0.0), (2, 1000, 3, 0.0), (2, 3152, 20, 0.0), (1, 3152, 20, 6.0),
                                            (2, 700, 4, -5.0)], ids=str)
def test_git_attention_vs_fp32_reference(n, n_vis, L, ramp):
    """Visual rows: the tcgen05 flash kernel (128-key chunks, tail chunk masked, query tiles that end inside a sample);
    text rows: the per-row-limit kernel.  ramp != 0 grows (or shrinks) the keys along the sequence so that the running
    maximum keeps moving by more than the lazy-rescale threshold and the O accumulators in TMEM are rescaled in later
    chunks (ramp > 0), or is set by the first chunk once and for all (ramp < 0)."""
    torch.manual_seed(n_vis + L)
    S = n_vis + L
    rows = n * S
    qkv = torch.randn(rows, 2304)
    if ramp != 0.0:
        pos = torch.cat([torch.arange(n_vis).repeat(n), n_vis + torch.arange(L).repeat(n)]).float() / S
        gain = 1.0 + abs(ramp) * (pos if ramp > 0 else 1.0 - pos)
        qkv[:, 768:1536] *= gain[:, None]
    qkv = qkv.to(torch.bfloat16).to(DEV)
    out = torch.empty(rows, 768, dtype=torch.bfloat16, device=DEV)
    _capi.check(_capi.lib().sasvqa_test_attention_git(qkv.data_ptr(), n, n_vis, L, out.data_ptr(),
                                                      torch.cuda.current_stream().cuda_stream), "attention_git")
    q_idx, k_idx = torch.arange(S)[:, None], torch.arange(S)[None, :]
    allowed = torch.where(q_idx < n_vis, k_idx < n_vis, k_idx <= q_idx).to(DEV)
    for s in range(n):
        idx = torch.cat([torch.arange(s * n_vis, (s + 1) * n_vis), torch.arange(n * n_vis + s * L, n * n_vis + (s + 1) * L)]).to(DEV)
        x = qkv[idx].float()
        q, k, v = (x[:, i * 768:(i + 1) * 768].view(S, 12, 64).transpose(0, 1) for i in range(3))
        sc = (q @ k.transpose(1, 2) / 8.0).masked_fill(~allowed, float("-inf"))
        want = (torch.softmax(sc, dim=-1) @ v).transpose(0, 1).reshape(S, 768)
        assert (out[idx].float() - want).abs().max().item() <= 2e-2


def test_vqa_hidden_and_logits_vs_oracle_and_hf_fixture(models, weights, golden_dir):
    enc, dec = models
    g = np.load(os.path.join(golden_dir, "git_vqa_hf.npz"))
    K = int(g["K"])
    frames = torch.stack([vit.image_processor_224(synth.make_clip(int(c), K)) for c in g["clip_ids"]])
    ids = torch.from_numpy(g["input_ids"])
    mask = torch.from_numpy(g["attention_mask"]).bool()
    oracle = git_oracle.GitVqaOracle(*weights)
    # text embeddings + projected visual tokens before any decoder block
    vis0, txt0 = vqa.vqa_hidden(frames, ids, enc, dec, n_layers=0)
    ref0, nv = oracle.hidden_states(frames, ids, n_layers=0)
    assert (txt0.cpu() - ref0[:, nv:]).abs().max().item() <= 2e-5
    assert torch.nn.functional.cosine_similarity(vis0.cpu(), ref0[:, :nv], dim=-1).min().item() >= 0.9995
    # after all six blocks
    vis6, txt6 = vqa.vqa_hidden(frames, ids, enc, dec, n_layers=dec.n_layers)
    ref6, _ = oracle.hidden_states(frames, ids)
    assert torch.nn.functional.cosine_similarity(vis6.cpu(), ref6[:, :nv], dim=-1).min().item() >= 0.999
    assert torch.nn.functional.cosine_similarity(txt6.cpu(), ref6[:, nv:], dim=-1)[mask].min().item() >= 0.999
    # logits of the text rows against HF's own (fixture) and the oracle
    logits = sas.vqa_logits(frames, ids, enc, dec).cpu()
    assert logits.shape == (2, ids.shape[1], dec.vocab)
    want = oracle(frames, ids)
    err = (logits - want).abs().amax(dim=-1)
    assert err[mask].max().item() <= 6e-2, err
    assert (logits[:, :, ::61] - torch.from_numpy(g["logits_probe"])).abs()[mask].max().item() <= 6e-2
    assert (logits[0, 3] - torch.from_numpy(g["logits_row"])).abs().max().item() <= 6e-2
    top5_idx, top5_val = torch.from_numpy(g["top5_idx"]), torch.from_numpy(g["top5_val"])
    excused = 0
    for b in range(2):
        for t in range(ids.shape[1]):
            if not mask[b, t]:
                continue
            if int(logits[b, t].argmax()) != int(top5_idx[b, t, 0]):
                assert float(top5_val[b, t, 0] - top5_val[b, t, 1]) <= 2 * float(err[b, t]), (b, t)
                excused += 1
    assert excused <= 1


def test_vqa_grouping_is_invisible_and_errors(models, weights):
    enc, dec = models
    frames = torch.stack([vit.image_processor_224(synth.make_clip(70 + c, 3)) for c in range(5)])     # S = 591 + 6
    ids = torch.randint(1000, synth.GIT_VOCAB, (5, 6), generator=torch.Generator().manual_seed(1))
    a = sas.vqa_logits(frames, ids, enc, dec)            # max_rows 4096 -> 6 samples per pass: one group
    small = vqa.GitDecoder(weights[2], max_rows=1300)    # 2 samples per pass: three groups
    try:
        b = sas.vqa_logits(frames, ids, enc, small)
        assert torch.equal(a, b)
        with pytest.raises(sas.SasvqaError):
            sas.vqa_logits(torch.zeros(1, 8, 3, 224, 224), ids[:1], enc, small)      # 8 * 197 + 6 rows > max_rows
    finally:
        small.close()
    with pytest.raises(ValueError):
        sas.vqa_logits(frames[:, 0], ids, enc, dec)                                   # rank 4
    with pytest.raises(sas.SasvqaError):
        sas.vqa_logits(frames, torch.zeros(5, 1025, dtype=torch.long), enc, dec)      # beyond the position table


def test_vqa_loss_vs_reference_expression_on_hf_logits(models, weights, golden_dir):
    """loss of MyGitForCausalLM.forward (modeling.py:208-215): fixture = the reference's expression on HF's logits."""
    enc, dec = models
    g = np.load(os.path.join(golden_dir, "git_vqa_hf.npz"))
    K = int(g["K"])
    frames = torch.stack([vit.image_processor_224(synth.make_clip(int(c), K)) for c in g["clip_ids"]])
    ids, labels = torch.from_numpy(g["input_ids"]), torch.from_numpy(g["labels"])
    loss, logits = vqa.vqa_loss(frames, ids, labels, enc, dec, want_logits=True)
    assert abs(float(loss) - float(g["loss"])) <= 3e-2, (float(loss), float(g["loss"]))
    assert torch.equal(logits, sas.vqa_logits(frames, ids, enc, dec))
    # the kernel's reduction against torch on the very same logits: fp32 rounding only
    want = torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, dec.vocab), labels[:, 1:].reshape(-1).to(DEV))
    assert abs(float(loss) - float(want)) <= 1e-4
    assert float(vqa.vqa_loss(frames, ids, labels, enc, dec)) == float(loss)              # scratch-logits path, same value
    small = vqa.GitDecoder(weights[2], max_rows=500)                                         # one sample per pass
    try:
        assert float(vqa.vqa_loss(frames, ids, labels, enc, small)) == float(loss)
    finally:
        small.close()
    assert torch.isnan(vqa.vqa_loss(frames, ids, torch.full_like(labels, -100), enc, dec))   # nothing to predict: 0 / 0, as torch


def test_vqa_generate_greedy_vs_fixture_and_teacher_forcing(models, weights, golden_dir):
    """Greedy decoding (modeling.py:333) with the visual keys/values cached per block: (1) every generated token is the
    argmax of the FULL forward on the prefix -- the cached path must agree bit for bit with sas.vqa_logits; (2) tokens
    equal the fixture's greedy chain on HF's uncached forward, a first mismatch excused only at a top-2 margin below
    2 x 6e-2 (after which the chains legitimately diverge)."""
    enc, dec = models
    g = np.load(os.path.join(golden_dir, "git_vqa_hf.npz"))
    K = int(g["K"])
    frames = torch.stack([vit.image_processor_224(synth.make_clip(int(c), K)) for c in g["clip_ids"]])
    prompt = torch.from_numpy(g["gen_prompt"])
    L0, Lmax = prompt.shape[1], g["gen_ids"].shape[1]
    out = vqa.vqa_generate(frames, prompt, enc, dec, max_length=Lmax, trim=False).cpu()
    assert out.shape == (2, Lmax) and torch.equal(out[:, :L0], prompt)
    logits = sas.vqa_logits(frames, out[:, :-1], enc, dec)                    # teacher forcing on our own output
    assert torch.equal(logits[:, L0 - 1:].argmax(dim=-1).cpu(), out[:, L0:])
    want, margins = torch.from_numpy(g["gen_ids"]), torch.from_numpy(g["gen_margins"])
    excused = 0
    for b in range(2):
        for t in range(L0, Lmax):
            if int(out[b, t]) != int(want[b, t]):
                assert float(margins[b, t - L0]) <= 2 * 6e-2, (b, t, out[b].tolist(), want[b].tolist())
                excused += 1
                break
    assert excused <= 1
    # eos handling: make the first generated token of sample 0 the eos id -> that sequence pads from then on, and with
    # trim the all-finished tail is cut where HF would stop
    eos = int(out[0, L0])
    cut = vqa.vqa_generate(frames[:1], prompt[:1], enc, dec, max_length=Lmax, eos_token_id=eos, pad_token_id=0).cpu()
    assert cut.shape == (1, L0 + 1) and int(cut[0, L0]) == eos
    both = vqa.vqa_generate(frames, prompt, enc, dec, max_length=Lmax, eos_token_id=eos, pad_token_id=0, trim=False).cpu()
    assert both[0, L0 + 1:].eq(0).all() and int(both[0, L0]) == eos
    if eos not in out[1, L0:].tolist():
        assert torch.equal(both[1], out[1])                                      # the other sequence is unaffected
    # every sequence finished at its first step: the group stops early (checked every 4th step) and the tail is still pad
    always = vqa.vqa_generate(frames[:1].repeat(2, 1, 1, 1, 1), prompt[:1].repeat(2, 1), enc, dec, max_length=40, eos_token_id=eos,
                              trim=False).cpu()
    assert always.shape == (2, 40) and always[:, L0].eq(eos).all() and always[:, L0 + 1:].eq(0).all()
    # grouping into passes is invisible
    small = vqa.GitDecoder(weights[2], max_rows=500)
    try:
        assert torch.equal(vqa.vqa_generate(frames, prompt, enc, small, max_length=Lmax, trim=False).cpu(), out)
    finally:
        small.close()
