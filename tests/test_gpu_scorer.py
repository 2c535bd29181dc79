"""GPU parity tests of the MIF caption cross-encoder (pytest -m gpu): the BERT sequence classifier of
gen_sample.py:79-88 through the C ABI, against the fp32 oracle (oracle/bert.py, pinned to HF's
BertForSequenceClassification by tests/golden/bert_scorer_hf.npz) on the same seeded weights and inputs.

Tolerances:
  embeddings + LayerNorm (fp32 arithmetic) ..... |d| <= 2e-5
  variable-length attention vs fp32 ............ |d| <= 2e-2 on bf16 outputs of O(1) values
  bf16 layers vs fp32 reference ................ final hidden state cosine >= 0.9995 per token
  logits ....................................... |d| <= 3e-2 (logit spread of the fixture ~1.3)
  selected caption indices ..................... identical except where the deciding scores differ by <= 2*eps,
                                                 eps = max |score_gpu - score_ref| of that QA sample
  host entry vs device entry, chunked vs one pass: bit-exact
"""
import os

import numpy as np
import pytest
import torch

from oracle import bert
import sasvqa_b200 as sas
from sasvqa_b200 import _capi, ops, synth
from sasvqa_b200.scorer import CaptionScorer, generate_inds

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
VOCAB = 2048


@pytest.fixture(scope="module")
def scorer_sd():
    return synth.random_scorer_state_dict(vocab=VOCAB)


@pytest.fixture(scope="module")
def scorer(scorer_sd):
    torch.cuda.set_device(0)
    sc = CaptionScorer(scorer_sd, max_tokens=2048)
    yield sc
    sc.close()


@pytest.fixture(scope="module")
def oracle_model(scorer_sd):
    return bert.BertScorerOracle(scorer_sd)


def _golden(golden_dir):
    return np.load(os.path.join(golden_dir, "bert_scorer_hf.npz"))


def _inputs(g, name):
    return (torch.from_numpy(g[f"{name}_input_ids"]), torch.from_numpy(g[f"{name}_token_type_ids"]),
            torch.from_numpy(g[f"{name}_attention_mask"]))


def test_gemm_erf_gelu_epilogue():
    torch.manual_seed(3)
    M, N, K = 777, 3072, 768
    a = (torch.randn(M, K) * 0.5).to(torch.bfloat16).to(DEV)
    b = (torch.randn(N, K) * 0.05).to(torch.bfloat16).to(DEV)
    bias = (torch.randn(N) * 0.1).to(DEV)
    want = torch.nn.functional.gelu(a.float() @ b.float().t() + bias)
    for use_simt in (True, False):
        got = ops.test_gemm(a, b, 4, bias, use_simt=use_simt).float()
        assert (got - want).abs().max().item() <= 2e-2 * max(1.0, want.abs().max().item())
    # the activation itself over its whole range, negative tail included: acc == x exactly (one non-zero product),
    # so the output must be the bf16 rounding of fp32 gelu(x) to within one bf16 ulp
    x = torch.cat([torch.linspace(-9.0, 9.0, 4001), torch.tensor([0.0, -0.0, 1e-6, -1e-6, 30.0, -30.0])])
    x = x.to(torch.bfloat16)
    a = torch.zeros(x.numel(), 64, dtype=torch.bfloat16)
    a[:, 0] = x
    b = torch.zeros(256, 64, dtype=torch.bfloat16)
    b[:, 0] = 1.0
    got = ops.test_gemm(a.to(DEV), b.to(DEV), 4, torch.zeros(256, device=DEV)).float().cpu()
    want = torch.nn.functional.gelu(x.double()).float()
    err = (got - want[:, None]).abs()
    assert (err <= want.abs()[:, None] * 2.0 ** -8 + 1e-9).all(), float((err - want.abs()[:, None] * 2.0 ** -8).max())


@pytest.mark.parametrize("lens", [[1], [2, 15, 16, 17], [64, 33, 48, 1, 20, 49] * 40, [64, 65, 63, 1, 128], [200, 3, 512, 77]], ids=str)
def test_attention_varlen_vs_fp32_reference(lens):
    torch.manual_seed(sum(lens))
    M = sum(lens)
    qkv = torch.randn(M, 2304).to(torch.bfloat16).to(DEV)
    cu = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32, device=DEV)
    out = torch.empty(M, 768, dtype=torch.bfloat16, device=DEV)
    _capi.check(_capi.lib().sasvqa_test_attention_varlen(qkv.data_ptr(), cu.data_ptr(), len(lens), max(lens), out.data_ptr(),
                                                         torch.cuda.current_stream().cuda_stream), "attention_varlen")
    r0 = 0
    for n in lens:
        x = qkv[r0:r0 + n].float()
        q, k, v = (x[:, i * 768:(i + 1) * 768].view(n, 12, 64).transpose(0, 1) for i in range(3))
        want = (torch.softmax(q @ k.transpose(1, 2) / 8.0, dim=-1) @ v).transpose(0, 1).reshape(n, 768)
        assert (out[r0:r0 + n].float() - want).abs().max().item() <= 2e-2
        r0 += n


def test_scorer_embeddings_and_hidden_vs_oracle_and_hf_fixture(scorer, oracle_model, golden_dir):
    g = _golden(golden_dir)
    ids, tts, msk = _inputs(g, "batch")
    lens = msk.sum(1).tolist()
    h0 = scorer.hidden(ids, tts, msk, n_layers=0).cpu()
    h12 = scorer.hidden(ids, tts, msk, n_layers=12).cpu()
    assert h0.shape[0] == sum(lens)
    # HF's own hidden states of the first pair
    assert (h0[:lens[0]] - torch.from_numpy(g["batch_hidden0_row0"])[:lens[0]]).abs().max().item() <= 2e-5
    ref0 = oracle_model.hidden_states(ids, tts, msk, n_layers=0)
    ref12 = oracle_model.hidden_states(ids, tts, msk, n_layers=12)
    r0 = 0
    for i, n in enumerate(lens):
        assert (h0[r0:r0 + n] - ref0[i, :n]).abs().max().item() <= 2e-5
        cos = torch.nn.functional.cosine_similarity(h12[r0:r0 + n], ref12[i, :n], dim=1)
        assert cos.min().item() >= 0.9995, (i, cos.min().item())
        r0 += n
    cos_hf = torch.nn.functional.cosine_similarity(h12[:lens[0]], torch.from_numpy(g["batch_hidden12_row0"])[:lens[0]], dim=1)
    assert cos_hf.min().item() >= 0.9995


def test_scorer_logits_and_indices_vs_hf_fixture(scorer, golden_dir):
    g = _golden(golden_dir)
    for name in ("batch", "edge"):                       # edge: shortest pair (3 tokens), 6 tokens, 512 tokens
        ids, tts, msk = _inputs(g, name)
        got = scorer(input_ids=ids, token_type_ids=tts, attention_mask=msk)[0].cpu()
        want = torch.from_numpy(g[f"{name}_logits"])
        assert (got - want).abs().max().item() <= 3e-2, (name, (got - want).abs().max().item())
    ids, tts, msk = _inputs(g, "batch")
    excused = 0
    for ds_rate in (1, 2):
        idx, scores = scorer.select_captions_host(ids, tts, msk, n_samples=3, K=3, ds_rate=ds_rate, want_scores=True)
        want_scores = torch.from_numpy(g["batch_logits"])[:, 0].view(3, 6)
        for s in range(3):
            eps = (scores[s] - want_scores[s]).abs().max().item()
            for a, b in zip(idx[s].tolist(), g[f"batch_inds_ds{ds_rate}"][s].tolist()):
                if a != b:
                    assert abs(float(want_scores[s, a]) - float(want_scores[s, b])) <= 2 * eps
                    excused += 1
    assert excused == 0, f"{excused} tie-excused picks on the fixture (expected none)"


def test_scorer_host_entry_chunking_and_batch_invariance(scorer, scorer_sd):
    tok = synth.SynthTokenizer(VOCAB)
    qa, caps = synth.make_qa_workload(12, 16, seed=7)
    text, pair = [], []
    for s in qa:
        text += [s["question"]] * 16
        pair += caps[f"video{s['video']}"]
    b = tok(text=text, text_pair=pair)
    n_tokens = int(b["attention_mask"].sum())
    assert n_tokens > scorer.max_tokens                   # 192 pairs do not fit one 2048-token pass: several groups
    dev_logits = scorer.logits(b["input_ids"], b["token_type_ids"], b["attention_mask"]).cpu()
    host_logits = scorer.logits_host(b["input_ids"], b["token_type_ids"], b["attention_mask"])
    assert torch.equal(dev_logits, host_logits)
    big = CaptionScorer(scorer_sd, max_tokens=16384)       # one pass
    try:
        assert torch.equal(big.logits(b["input_ids"], b["token_type_ids"], b["attention_mask"]).cpu(), dev_logits)
        # a pair's logits do not depend on what it is batched or padded with (the reference pads per QA sample)
        one = tok(text=text[16:32], text_pair=pair[16:32])
        assert torch.equal(big.logits_host(one["input_ids"], one["token_type_ids"], one["attention_mask"]), dev_logits[16:32])
    finally:
        big.close()
    idx, scores = scorer.select_captions_host(b["input_ids"], b["token_type_ids"], b["attention_mask"], 12, K=4, ds_rate=2,
                                              want_scores=True)
    assert torch.equal(scores.view(-1), dev_logits[:, 0])
    for s in range(12):
        want = bert.mif_indices_from_logits(dev_logits[s * 16:(s + 1) * 16], 4, 2)
        assert idx[s].tolist() == want


def test_generate_inds_mirror_vs_oracle_loop(scorer, oracle_model):
    tok = synth.SynthTokenizer(VOCAB)
    qa, caps = synth.make_qa_workload(10, 8, seed=11)
    qa = qa + [dict(qa[3], question="how many people are playing ?")]       # two questions on one video
    got = generate_inds(tok, scorer, qa, caps, K=3, ds_rate=1, dataset="msvd_qa", samples_per_call=4)
    want = bert.generate_inds(tok, oracle_model, qa, caps, K=3, ds_rate=1)
    assert [r["question"] for r in got] == [r["question"] for r in want]
    excused = 0
    for sample, r_got, r_want in zip(qa, got, want):
        if r_got["sampled_inds"] == r_want["sampled_inds"]:
            continue
        captions = caps[f"video{sample['video']}"]
        b = tok(text=[sample["question"]] * len(captions), text_pair=captions)
        ref = oracle_model(b["input_ids"], b["token_type_ids"], b["attention_mask"])[:, 0]
        mine = scorer.logits_host(b["input_ids"], b["token_type_ids"], b["attention_mask"])[:, 0]
        eps = (ref - mine).abs().max().item()
        for a, c in zip(r_got["sampled_inds"], r_want["sampled_inds"]):
            if a != c:
                assert abs(float(ref[a]) - float(ref[c])) <= 2 * eps, (r_got, r_want, eps)
                excused += 1
    assert excused <= 1, excused


def test_scorer_errors(scorer):
    ids = torch.randint(1000, VOCAB, (2, 8))
    with pytest.raises(ValueError):
        scorer.logits(ids, None, torch.tensor([[0, 1, 1, 1, 1, 1, 1, 1], [1] * 8]))          # left padding
    with pytest.raises(sas.SasvqaError):
        scorer.logits_host(ids, None, torch.tensor([[1, 0, 1, 1, 1, 1, 1, 1], [1] * 8]))       # hole in the mask
    bad = ids.clone()
    bad[1, 2] = VOCAB
    with pytest.raises(sas.SasvqaError):
        scorer.logits_host(bad, None, None)
    with pytest.raises(RuntimeError):                                                           # torch.topk's error
        scorer.select_captions_host(ids, None, None, n_samples=1, K=3, ds_rate=1)
    with pytest.raises(sas.SasvqaError):
        scorer.logits(torch.zeros(1, 513, dtype=torch.long)) if False else scorer.logits_host(
            torch.zeros(1, 513, dtype=torch.long), None, None)
    with pytest.raises(sas.SasvqaError):
        CaptionScorer({k: v for k, v in synth.random_scorer_state_dict(vocab=64).items()}, max_tokens=100)
