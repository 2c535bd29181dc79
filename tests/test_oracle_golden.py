"""Pins the CPU oracle: against the fixtures produced by the reference itself
(oracle/make_golden.py) and, when /root/reference is present, against the live reference."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import mdf, ref_loader, vit
import sasvqa_b200.synth as synth


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _iter_mdf_cases(g):
    D = int(g["D"])
    fo = io = 0
    for T, K, W, failure, zeros, err, n_idx in g["meta"].tolist():
        feats = g["feats"][fo:fo + T * D].reshape(T, D)
        idx = g["indices"][io:io + n_idx].tolist()
        fo += T * D
        io += n_idx
        yield T, K, W, failure, zeros, err, feats, idx


def assert_same_or_tied(got, want, scores):
    """Index lists must match position by position except where both picks carry EXACTLY
    the same score (torch.topk's order among equal values is implementation-defined)."""
    assert len(got) == len(want)
    for a, b in zip(got, want):
        if a != b:
            assert float(scores[a]) == float(scores[b]), (got, want)


def test_mdf_select_matches_reference_fixtures(golden_dir):
    g = _load(golden_dir, "mdf_select.npz")
    n = n_fb = n_err = 0
    for T, K, W, failure, zeros, err, feats, want in _iter_mdf_cases(g):
        n += 1
        if T == 0:
            assert zeros == 1
            continue
        f = torch.nn.functional.normalize(torch.from_numpy(feats.copy()))
        if err:
            n_err += 1
            with pytest.raises(RuntimeError):
                mdf.mdf_indices_from_feats(f, K, W)
            continue
        idx, status, lcl, _ = mdf.mdf_indices_from_feats(f, K, W)
        assert status == (mdf.STATUS_FALLBACK if failure else mdf.STATUS_OK)
        n_fb += failure
        assert_same_or_tied(idx, want, lcl)
    assert n >= 100 and n_fb > 10 and n_err > 5


def test_mdf_select_exact_where_no_ties(golden_dir):
    """On cases whose interior scores are all distinct and whose picks avoid the zero
    borders the oracle must be identical to the reference with no excuses."""
    g = _load(golden_dir, "mdf_select.npz")
    exact = 0
    for T, K, W, failure, zeros, err, feats, want in _iter_mdf_cases(g):
        if T == 0 or err:
            continue
        f = torch.nn.functional.normalize(torch.from_numpy(feats.copy()))
        idx, status, lcl, _ = mdf.mdf_indices_from_feats(f, K, W)
        vals = lcl.numpy()
        if len(np.unique(vals[want])) == len(want) and np.all(vals[want] != 0):
            assert idx == want
            exact += 1
    assert exact >= 40


def test_uniform_git6_mif_fixtures(golden_dir):
    g = _load(golden_dir, "samplers_misc.npz")
    o = 0
    for T, K in g["uni_meta"].tolist():
        assert mdf.uniform_indices(T, K) == g["uni_idx"][o:o + K].tolist()
        o += K
    o = 0
    rng = np.random.RandomState(666)
    for T, K in g["git_meta"].tolist():
        got = mdf.git6_indices(T, K, 4, rng=rng)
        assert got.tolist() == g["git_idx"][o:o + K].tolist()
        o += K
    o = so = 0
    for T, K, ds in g["mif_meta"].tolist():
        s = g["mif_scores"][so:so + T]
        assert mdf.mif_select(s, K, ds) == g["mif_idx"][o:o + K].tolist()
        o += K
        so += T


def test_known_answer_edges():
    # T <= 2W: all-zero scores, argmax 0
    lcl = mdf.local_average(torch.eye(10), 8)
    assert float(lcl.abs().sum()) == 0.0
    # W = -1 with T < 20 -> W = 0 -> every score (0 - 1) / (0 - 1) = 1
    f = torch.nn.functional.normalize(torch.randn(12, 8))
    idx, status, lcl, _ = mdf.mdf_indices_from_feats(f, 4, -1)
    assert torch.all(lcl == 1.0)
    # two equal-height peaks: heap tie falls to the smaller left edge
    lcl = torch.zeros(64)
    lcl[30] = 0.9
    lcl[10] = 0.5
    lcl[50] = 0.5
    assert mdf.greedy_select(lcl, 3, 4) == [30, 10, 50]
    # spacing: picks differ by >= W
    lcl = torch.rand(200)
    picks = mdf.greedy_select(lcl, 10, 8)
    assert len(picks) == 10 and min(abs(a - b) for i, a in enumerate(picks) for b in picks[:i]) >= 8
    # fallback with T < K raises like torch.topk
    with pytest.raises(RuntimeError):
        mdf.mdf_select(torch.zeros(4), 8, 8)


def test_encoder_restatement_matches_hf_fixture(golden_dir):
    g = _load(golden_dir, "encoder_hf.npz")
    sd = synth.random_encoder_state_dict(int(g["seed"]))
    u8 = synth.make_clip(int(g["clip_id"]), int(g["T"]))
    px = vit.image_processor_224(u8)
    assert np.abs(px[:, :, ::37, ::41].numpy() - g["pixel_probe"]).max() < 1e-6
    assert np.abs(px.double().sum(dim=(1, 2, 3)).numpy() - g["pixel_sum"]).max() < 1e-1
    enc = vit.VitOracle(sd)
    hid = enc.forward_hidden(px)
    assert np.abs(hid[:, ::49, ::64].numpy() - g["hidden_probe"]).max() < 2e-4
    feats = torch.nn.functional.normalize(hid.mean(dim=1)).numpy()
    assert np.abs(feats - g["feats"]).max() < 1e-5


def test_e2e_fixture_selection_from_reference_features(golden_dir):
    g = _load(golden_dir, "mdf_e2e_hf.npz")
    for tag in ("c1", "t64k8w4", "t128k8w8"):
        cid, T, K, W, failure = g[tag + "_meta"].tolist()
        feats = torch.from_numpy(g[tag + "_feats"])
        idx, status, lcl, _ = mdf.mdf_indices_from_feats(feats, K, W)
        assert status == failure
        assert np.abs(lcl.numpy() - g[tag + "_lcl"]).max() < 1e-6
        assert idx == g[tag + "_indices"].tolist()


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference absent (GPU box)")
def test_live_reference_agrees_on_fresh_random_cases():
    ref_mdf, ref_uni = ref_loader.load_sampler_fns()
    rng = np.random.RandomState(4242)
    checked = 0
    for trial in range(120):
        T = int(rng.choice([7, 24, 50, 64, 96, 128, 200, 333]))
        K = int(rng.choice([2, 4, 8, 16]))
        W = int(rng.choice([-1, 1, 2, 4, 8]))
        raw = torch.from_numpy(np.cumsum(rng.randn(T, 6).astype(np.float32), axis=0))

        class Stub:
            def __call__(self, frames):
                i = frames.reshape(frames.shape[0], -1)[:, 0].long()
                return SimpleNamespace(pooler_output=raw[i])

        frames = torch.arange(T, dtype=torch.float32).view(T, 1, 1, 1)
        dc = {"Failure": 0, "Zeros": 0}
        try:
            want = ref_mdf(frames, Stub(), K, W, dc).reshape(-1).long().tolist()
        except RuntimeError:
            with pytest.raises(RuntimeError):
                mdf.sample_representative_frames(frames, Stub(), K, W, {"Failure": 0, "Zeros": 0})
            continue
        dc2 = {"Failure": 0, "Zeros": 0}
        out, aux = mdf.sample_representative_frames(frames, Stub(), K, W, dc2, return_aux=True)
        assert dc2 == dc
        assert_same_or_tied(aux["indices"], want, aux["lcl_avg"])
        assert out.reshape(-1).long().tolist() == aux["indices"]
        checked += 1
    assert checked >= 80
    frames = torch.arange(100, dtype=torch.float32).view(100, 1)
    assert ref_uni(frames, K=8).reshape(-1).long().tolist() == mdf.uniform_indices(100, 8)


# ------------------------------------------------------------------ K0 oracle: resize + centre crop
def test_resize_restatement_matches_hf_processor_fixture(golden_dir):
    """oracle/resize.py == the installed HF CLIPImageProcessor (fixture written by oracle/make_golden.py)."""
    from oracle import resize
    g = _load(golden_dir, "resize_hf.npz")
    assert [tuple(c) for c in g["cases"].tolist()] == resize.RESIZE_CASES
    for h, w in resize.RESIZE_CASES:
        frames = resize.resize_case_frames(h, w)
        got = resize.resize_crop_u8(frames)
        assert got.shape == (2, 224, 224, 3) and got.dtype == np.uint8
        assert np.array_equal(got, g[f"out_{h}x{w}"]), (h, w)
        px = vit.image_processor_224(torch.from_numpy(got))            # rescale + normalise of the cropped frame
        assert np.abs(px[:, :, ::37, ::41].numpy() - g[f"pixel_probe_{h}x{w}"]).max() <= 1e-6


def test_resize_restatement_matches_torch_cpu_kernel_live():
    """Live pin against the third-party code itself: torch's CPU uint8 anti-aliased bicubic resampler
    (what torchvision's resize calls for the reference's host image processor)."""
    from oracle import resize
    for h, w in [(250, 333), (224, 300), (500, 224), (97, 131), (448, 448)]:
        rng = np.random.RandomState(h + w)
        img = rng.randint(0, 256, (1, h, w, 3), dtype=np.uint8)
        oh, ow = resize.output_size(h, w)
        ref = torch.nn.functional.interpolate(torch.from_numpy(img).permute(0, 3, 1, 2), size=(oh, ow), mode="bicubic",
                                              antialias=True, align_corners=False).permute(0, 2, 3, 1).numpy()
        top, left = int((oh - 224) / 2.0), int((ow - 224) / 2.0)
        assert np.array_equal(resize.resize_crop_u8(img), ref[:, top:top + 224, left:left + 224]), (h, w)
    same = np.random.RandomState(1).randint(0, 256, (2, 224, 224, 3), dtype=np.uint8)
    assert np.array_equal(resize.resize_crop_u8(same), same)


# ------------------------------------------------------------------ row f2 oracle: downstream visual tokens
def test_visual_tokens_restatement_matches_hf_fixture(golden_dir):
    """oracle.vit.visual_tokens == HF GitVisionModel + HF GitProjection as src/modeling/modeling.py:76-95 chains them."""
    g = _load(golden_dir, "visual_tokens_hf.npz")
    sd, psd = synth.random_encoder_state_dict(synth.REF_SEED), synth.random_projection_state_dict()
    frames = torch.stack([vit.image_processor_224(synth.make_clip(int(c), int(g["T"]))) for c in g["clip_ids"]])
    enc = vit.VitOracle(sd)
    hidden = vit.visual_tokens(frames, enc, None)
    tokens = vit.visual_tokens(frames, enc, psd)
    assert tokens.shape == (2, 2 * 197, 768)
    assert np.abs(hidden[:, ::29, ::48].numpy() - g["hidden_probe"]).max() <= 2e-4
    assert np.abs(tokens[:, ::29, ::48].numpy() - g["tokens_probe"]).max() <= 2e-4
    assert np.abs(tokens.double().sum(dim=-1).numpy() - g["tokens_rowsum"]).max() <= 2e-2


def test_bert_scorer_restatement_matches_hf_fixture(golden_dir):
    """oracle/bert.py against HF BertForSequenceClassification's own outputs (tests/golden/bert_scorer_hf.npz,
    written by oracle/make_golden.py with the seeded random bert-base weights), and the reference's MIF
    expression (gen_sample.py:83-88) evaluated on HF's logits."""
    from oracle import bert
    import sasvqa_b200.synth as synth
    g = _load(golden_dir, "bert_scorer_hf.npz")
    sd = synth.random_scorer_state_dict(vocab=int(g["vocab"]))
    model = bert.BertScorerOracle(sd)
    for name in ("batch", "edge"):
        ids, tts, msk = (torch.from_numpy(g[f"{name}_{k}"]) for k in ("input_ids", "token_type_ids", "attention_mask"))
        logits = model(ids, tts, msk)
        assert (logits - torch.from_numpy(g[f"{name}_logits"])).abs().max().item() <= 1e-4
        if name == "batch":
            h0 = model.hidden_states(ids, tts, msk, n_layers=0)[0]
            h12 = model.hidden_states(ids, tts, msk)[0]
            n = int(msk[0].sum())
            assert (h0[:n] - torch.from_numpy(g["batch_hidden0_row0"])[:n]).abs().max().item() <= 1e-5
            assert (h12[:n] - torch.from_numpy(g["batch_hidden12_row0"])[:n]).abs().max().item() <= 1e-4
            for ds_rate in (1, 2):
                for s in range(3):
                    got = bert.mif_indices_from_logits(logits[s * 6:(s + 1) * 6], 3, ds_rate)
                    assert got == g[f"batch_inds_ds{ds_rate}"][s].tolist()


def test_bert_scorer_padding_is_invisible_in_the_oracle():
    """The property the packed GPU path relies on: a pair's logits do not depend on how far it is padded."""
    from oracle import bert
    import sasvqa_b200.synth as synth
    sd = synth.random_scorer_state_dict(vocab=512)
    model = bert.BertScorerOracle(sd)
    tok = synth.SynthTokenizer(512)
    short = tok(text=["what is the man doing ?"], text_pair=["a man is cooking"])
    padded = tok(text=["what is the man doing ?", "what is the man doing ?"],
                 text_pair=["a man is cooking", "a woman is riding a horse near the water with two people"])
    a = model(short["input_ids"], short["token_type_ids"], short["attention_mask"])
    b = model(padded["input_ids"], padded["token_type_ids"], padded["attention_mask"])
    assert (a[0] - b[0]).abs().max().item() <= 2e-5


def test_git_vqa_restatement_matches_hf_fixture(golden_dir):
    """oracle/git.py (MyGitForCausalLM's inference forward, src/modeling/modeling.py:29-232) against HF GitForCausalLM's
    own logits with the temporal embeddings zeroed (tests/golden/git_vqa_hf.npz)."""
    from oracle import git as git_oracle, vit
    import sasvqa_b200.synth as synth
    g = _load(golden_dir, "git_vqa_hf.npz")
    K = int(g["K"])
    frames = torch.stack([vit.image_processor_224(synth.make_clip(int(c), K)) for c in g["clip_ids"]])
    ids = torch.from_numpy(g["input_ids"])
    mask = torch.from_numpy(g["attention_mask"]).bool()
    model = git_oracle.GitVqaOracle(synth.random_encoder_state_dict(synth.REF_SEED), synth.random_projection_state_dict(),
                                    synth.random_git_decoder_state_dict())
    logits = model(frames, ids)
    assert (logits[:, :, ::61] - torch.from_numpy(g["logits_probe"])).abs()[mask].max().item() <= 1e-4
    assert (logits[0, 3] - torch.from_numpy(g["logits_row"])).abs().max().item() <= 1e-4
    assert torch.equal(logits.topk(1, dim=-1).indices[..., 0][mask], torch.from_numpy(g["top5_idx"])[..., 0][mask])
    # greedy decoding (modeling.py:333) against the chain of argmaxes of HF's own uncached forward
    gen, _ = model.generate(frames, torch.from_numpy(g["gen_prompt"]), max_length=g["gen_ids"].shape[1])
    assert torch.equal(gen, torch.from_numpy(g["gen_ids"]))
    # the reference's loss expression (modeling.py:208-215) on HF's logits
    assert abs(float(model.loss(frames, ids, torch.from_numpy(g["labels"]))) - float(g["loss"])) <= 1e-4
