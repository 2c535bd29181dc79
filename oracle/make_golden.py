"""Generates tests/golden/*.npz by running the REFERENCE ITSELF in the build container.
TEST INFRASTRUCTURE ONLY.  Run:  python -m oracle.make_golden  (needs /root/reference).

The reference has no tests or golden vectors, so these fixtures are
produced from its own functions:

* ``sample_representative_frames`` / ``sample_frames_uniform``
  (/root/reference/src/preprocessing/datautils/utils.py:31,96) imported unmodified;
* ``sample_frame_indices`` (extract_features.py:32-39) compiled from its AST node;
* the MIF expression ``scores[::ds_rate].topk(K)[1]`` then ``i * ds_rate``
  (gen_sample.py:87-88) evaluated verbatim with CPU torch (the module itself cannot be
  imported: h5py / tensorboardX are absent);
* HF ``GitVisionModel`` / ``CLIPImageProcessor`` (the third-party code the reference calls)
  with the repo's seeded random ViT-B/16 weights.
"""
from __future__ import annotations

import os
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import mdf, ref_loader, vit  # noqa: E402
import sasvqa_b200.synth as synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


class StubEncoder:
    """Returns precomputed pooled rows; frames carry their own index in element 0."""

    def __init__(self, raw):
        self.raw = raw

    def __call__(self, frames):
        idx = frames.reshape(frames.shape[0], -1)[:, 0].long()
        return SimpleNamespace(pooler_output=self.raw[idx])


def structured_feats(rng: np.random.RandomState, T: int, D: int, kind: int) -> np.ndarray:
    if kind == 0:      # iid
        return rng.randn(T, D).astype(np.float32)
    if kind == 1:      # random walk (smooth drift)
        return np.cumsum(rng.randn(T, D).astype(np.float32) * 0.3, axis=0) + rng.randn(1, D).astype(np.float32)
    # scenes + noise
    n_sc = max(2, T // 16)
    centers = rng.randn(n_sc, D).astype(np.float32)
    cuts = np.sort(rng.choice(np.arange(1, max(T, 2)), size=min(n_sc - 1, max(T - 1, 1)), replace=False)) if T > 1 else []
    sc = np.searchsorted(cuts, np.arange(T), side="right") if T > 1 else np.zeros(T, int)
    return (centers[sc] + 0.25 * rng.randn(T, D)).astype(np.float32)


def gen_mdf_select():
    ref_mdf, _ = ref_loader.load_sampler_fns()
    rng = np.random.RandomState(666)
    D = 8
    Ts = [0, 1, 2, 5, 15, 16, 17, 20, 21, 33, 40, 64, 100, 128, 129, 200, 257, 300, 512, 600]
    cases = []
    for T in Ts:
        for K, W in [(16, 8), (8, 8), (8, 4), (4, 2), (32, 8), (16, -1), (1, 8), (3, 1), (16, 16), (5, 0), (32, 3)]:
            cases.append((T, K, W, len(cases) % 3))
    feats_all, meta, idx_all = [], [], []
    for (T, K, W, kind) in cases:
        raw = torch.from_numpy(structured_feats(rng, T, D, kind)) if T > 0 else torch.zeros(0, D)
        frames = torch.arange(T, dtype=torch.float32).view(T, 1, 1, 1)
        dc = {"Failure": 0, "Zeros": 0}
        err = 0
        try:
            out = ref_mdf(frames, StubEncoder(raw), K, W, dc)
            idx = [] if dc["Zeros"] else [int(v) for v in out.reshape(out.shape[0], -1)[:, 0].tolist()]
            if dc["Zeros"]:
                assert tuple(out.shape) == (K, 3, 224, 224) and float(out.abs().sum()) == 0.0
        except RuntimeError:
            err, idx = 1, []
        meta.append([T, K, W, dc["Failure"], dc["Zeros"], err, len(idx)])
        feats_all.append(raw.numpy().reshape(-1))
        idx_all.append(np.asarray(idx, dtype=np.int64))
    np.savez_compressed(
        os.path.join(GOLD, "mdf_select.npz"),
        meta=np.asarray(meta, dtype=np.int64), D=np.int64(D),
        feats=np.concatenate(feats_all).astype(np.float32),
        indices=np.concatenate(idx_all) if idx_all else np.zeros(0, np.int64))
    print("mdf_select:", len(cases), "cases; fallback", sum(m[3] for m in meta), "errors", sum(m[5] for m in meta))


def gen_misc():
    _, ref_uniform = ref_loader.load_sampler_fns()
    ref_git6 = ref_loader.load_sample_frame_indices()
    uni_meta, uni_idx = [], []
    for T in [8, 9, 16, 17, 31, 64, 100, 128, 300, 512, 1000]:
        for K in [1, 3, 6, 8, 16, 32]:
            if K > T:
                continue
            frames = torch.arange(T, dtype=torch.float32).view(T, 1)
            out = ref_uniform(frames, K=K)
            uni_meta.append([T, K])
            uni_idx.append(out.reshape(-1).long().numpy())
    git_meta, git_idx = [], []
    np.random.seed(666)      # extract_features.py:140
    for T in [33, 64, 65, 100, 128, 300, 512]:
        for K in [6, 8, 16]:
            if T <= 4 * K:
                continue
            frames = torch.arange(T, dtype=torch.float32)
            out = ref_git6(frames, K, 4, T)
            git_meta.append([T, K])
            git_idx.append(out.long().numpy())
    rng = np.random.RandomState(667)
    mif_meta, mif_scores, mif_idx = [], [], []
    for T in [8, 16, 17, 64, 128, 512]:
        for K, ds_rate in [(8, 1), (8, 2), (4, 1), (6, 2), (3, 3), (16, 1)]:
            if K > len(range(0, T, ds_rate)):
                continue
            s = torch.from_numpy(rng.randn(T).astype(np.float32))
            inds = s[::ds_rate].topk(K)[1].detach().cpu().tolist()     # gen_sample.py:87
            inds = [i * ds_rate for i in inds]                         # gen_sample.py:88
            mif_meta.append([T, K, ds_rate])
            mif_scores.append(s.numpy())
            mif_idx.append(np.asarray(inds, dtype=np.int64))
    np.savez_compressed(
        os.path.join(GOLD, "samplers_misc.npz"),
        uni_meta=np.asarray(uni_meta), uni_idx=np.concatenate(uni_idx),
        git_meta=np.asarray(git_meta), git_idx=np.concatenate(git_idx),
        mif_meta=np.asarray(mif_meta), mif_scores=np.concatenate(mif_scores), mif_idx=np.concatenate(mif_idx))
    print("misc: uniform", len(uni_meta), "git6", len(git_meta), "mif", len(mif_meta))


def recover_indices(frames, out):
    flat = frames.reshape(frames.shape[0], -1)
    res = []
    for k in range(out.shape[0]):
        hit = (flat == out[k].reshape(1, -1)).all(dim=1).nonzero().reshape(-1)
        assert hit.numel() == 1
        res.append(int(hit[0]))
    return res


def gen_encoder_and_e2e():
    from transformers import CLIPImageProcessor
    sd = synth.random_encoder_state_dict(synth.REF_SEED)
    hf = vit.hf_model_from_state_dict(sd)
    # --- image processor + encoder features on 4 frames
    u8 = synth.make_clip(0, 4)
    proc = CLIPImageProcessor()
    px_hf = torch.from_numpy(np.stack(proc(images=[f.numpy() for f in u8])["pixel_values"]))
    px = vit.image_processor_224(u8)
    print("image processor max |diff| vs restatement:", float((px_hf - px).abs().max()))
    with torch.no_grad():
        hid = hf(px_hf).last_hidden_state
    feats = torch.nn.functional.normalize(hid.mean(dim=1))
    np.savez_compressed(
        os.path.join(GOLD, "encoder_hf.npz"),
        seed=np.int64(synth.REF_SEED), clip_id=np.int64(0), T=np.int64(4),
        pixel_probe=px_hf[:, :, ::37, ::41].numpy(), pixel_sum=px_hf.double().sum(dim=(1, 2, 3)).numpy(),
        hidden_probe=hid[:, ::49, ::64].numpy(), feats=feats.numpy())
    # --- end-to-end MDF through the reference function + HF encoder
    ref_mdf, _ = ref_loader.load_sampler_fns()
    out = {}
    for tag, (cid, T, K, W) in {"c1": (1, 64, 16, 8), "t64k8w4": (1, 64, 8, 4), "t128k8w8": (2, 128, 8, 8)}.items():
        frames = vit.image_processor_224(synth.make_clip(cid, T))
        dc = {"Failure": 0, "Zeros": 0}
        t0 = time.time()
        with torch.no_grad():
            sel = ref_mdf(frames, hf, K, W, dc)
        dt = time.time() - t0
        idx = recover_indices(frames, sel)
        with torch.no_grad():
            _, aux = mdf.sample_representative_frames(frames, hf, K, W, {"Failure": 0, "Zeros": 0}, return_aux=True)
        assert aux["indices"] == idx or all(
            float(aux["lcl_avg"][a]) == float(aux["lcl_avg"][b]) for a, b in zip(aux["indices"], idx)), (aux["indices"], idx)
        out[tag + "_meta"] = np.asarray([cid, T, K, W, dc["Failure"]], dtype=np.int64)
        out[tag + "_indices"] = np.asarray(idx, dtype=np.int64)
        out[tag + "_lcl"] = aux["lcl_avg"].numpy()
        out[tag + "_feats"] = aux["feats"].numpy()
        print(f"e2e {tag}: T={T} K={K} W={W} failure={dc['Failure']} idx={idx} ({dt:.1f}s)")
    np.savez_compressed(os.path.join(GOLD, "mdf_e2e_hf.npz"), **out)


def gen_visual_tokens():
    """HF GitVisionModel + HF GitProjection (the modules the reference's MyGitModel.forward calls,
    src/modeling/modeling.py:76-95) on 2 clips x 2 frames with the seeded weights."""
    from transformers import GitConfig
    from transformers.models.git.modeling_git import GitProjection
    sd = synth.random_encoder_state_dict(synth.REF_SEED)
    psd = synth.random_projection_state_dict()
    hf = vit.hf_model_from_state_dict(sd)
    proj = GitProjection(GitConfig()).eval()
    proj.load_state_dict(psd)
    frames = torch.stack([vit.image_processor_224(synth.make_clip(c, 2)) for c in (50, 51)])      # [2, 2, 3, 224, 224]
    with torch.no_grad():
        hidden = torch.cat([hf(frames[:, f]).last_hidden_state for f in range(frames.shape[1])], dim=1)
        tokens = proj(hidden)
        mine = vit.visual_tokens(frames, vit.VitOracle(sd), psd)
    print("visual tokens: restatement vs HF max |diff|:", float((mine - tokens).abs().max()))
    np.savez_compressed(os.path.join(GOLD, "visual_tokens_hf.npz"), clip_ids=np.asarray([50, 51]), T=np.int64(2),
                        hidden_probe=hidden[:, ::29, ::48].numpy(), tokens_probe=tokens[:, ::29, ::48].numpy(),
                        tokens_rowsum=tokens.double().sum(dim=-1).numpy())


def gen_resize():
    """The installed HF CLIPImageProcessor (the reference's `self.processor`, prefetch_loader.py:74) on frames
    that need the shortest-edge bicubic resize + centre crop.  Its fp32 output is inverted to the uint8 image
    it normalised (exactly recoverable: the map u -> (u - 255 mean) / (255 std) is injective on 0..255)."""
    from transformers import CLIPImageProcessor
    from oracle import resize
    from oracle.resize import RESIZE_CASES, resize_case_frames
    proc = CLIPImageProcessor()
    mean = np.asarray(vit.IMAGE_MEAN, dtype=np.float64).reshape(1, 3, 1, 1)
    std = np.asarray(vit.IMAGE_STD, dtype=np.float64).reshape(1, 3, 1, 1)
    out = {"cases": np.asarray(RESIZE_CASES, dtype=np.int64)}
    for h, w in RESIZE_CASES:
        frames = resize_case_frames(h, w)
        px = np.stack(proc(images=[f for f in frames])["pixel_values"])          # [2, 3, 224, 224] fp32
        u = np.rint((px.astype(np.float64) * std + mean) * 255.0)
        assert np.abs((u / 255.0 - mean) / std - px).max() < 1e-5
        u8 = u.astype(np.uint8).transpose(0, 2, 3, 1)                              # [2, 224, 224, 3]
        mine = resize.resize_crop_u8(frames)
        print(f"resize {h}x{w}: restatement == HF processor: {np.array_equal(mine, u8)}")
        out[f"out_{h}x{w}"] = u8
        out[f"pixel_probe_{h}x{w}"] = px[:, :, ::37, ::41]
    np.savez_compressed(os.path.join(GOLD, "resize_hf.npz"), **out)


SCORER_VOCAB = 2048          # the vocabulary is only a gather table: a small one keeps weight generation fast


def scorer_golden_inputs():
    """Tokenized (question, caption) pairs the way gen_sample.py:80 builds them, plus edge rows: the shortest
    possible pair, a single-text row, and a 512-token pair (truncation)."""
    tok = synth.SynthTokenizer(SCORER_VOCAB)
    qa, caps = synth.make_qa_workload(3, 6, seed=synth.REF_SEED)
    text, pair = [], []
    for s in qa:
        text += [s["question"]] * 6
        pair += caps[f"video{s['video']}"]
    batch = tok(text=text, text_pair=pair, padding=True, truncation=True, return_tensors="pt")
    long_q = " ".join(synth._WORDS[i % len(synth._WORDS)] for i in range(300))
    long_c = " ".join(synth._WORDS[(7 * i) % len(synth._WORDS)] for i in range(400))
    edge = tok(text=["", "what", long_q], text_pair=["", "a dog", long_c], padding=True, truncation=True, return_tensors="pt")
    return qa, caps, batch, edge


def gen_bert_scorer():
    """HF BertForSequenceClassification (the class gen_sample.py:160 instantiates for a BERT checkpoint) with the
    repo's seeded random bert-base weights -> logits of the golden inputs, and the reference's MIF expression
    (gen_sample.py:83-88) evaluated on them."""
    from transformers import BertConfig, BertForSequenceClassification
    from oracle import bert
    sd = synth.random_scorer_state_dict(vocab=SCORER_VOCAB)
    model = BertForSequenceClassification(BertConfig(vocab_size=SCORER_VOCAB, num_labels=synth.BERT_LABELS)).eval()
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in k or "token_type_ids" in k for k in missing), (missing, unexpected)
    qa, caps, batch, edge = scorer_golden_inputs()
    mine = bert.BertScorerOracle(sd)
    out = {}
    with torch.no_grad():
        for name, b in (("batch", batch), ("edge", edge)):
            ref = model(**b)[0]
            got = mine(b["input_ids"], b["token_type_ids"], b["attention_mask"])
            print(f"bert scorer {name}: shape {tuple(ref.shape)}, restatement max |d logit| = {(ref - got).abs().max():.2e}")
            out[f"{name}_input_ids"] = b["input_ids"].numpy()
            out[f"{name}_token_type_ids"] = b["token_type_ids"].numpy()
            out[f"{name}_attention_mask"] = b["attention_mask"].numpy()
            out[f"{name}_logits"] = ref.numpy()
        hs = model.bert(**batch, output_hidden_states=True).hidden_states
        out["batch_hidden0_row0"] = hs[0][0].numpy()              # embeddings output of the first pair
        out["batch_hidden12_row0"] = hs[12][0].numpy()
        # gen_sample.py:83-88 verbatim per QA sample (K = 3; ds_rate 1 and 2)
        logits = torch.from_numpy(out["batch_logits"])
        for ds_rate in (1, 2):
            rows = []
            for g in range(3):
                scores = logits[g * 6:(g + 1) * 6][:, 0]
                inds = scores[::ds_rate].topk(3)[1].detach().cpu().tolist()
                rows.append([i * ds_rate for i in inds])
            out[f"batch_inds_ds{ds_rate}"] = np.asarray(rows, dtype=np.int64)
    np.savez_compressed(os.path.join(GOLD, "bert_scorer_hf.npz"), vocab=np.int64(SCORER_VOCAB), **out)


def gen_git_vqa():
    """HF GitForCausalLM (the class the reference's MyGitForCausalLM subclasses, src/modeling/modeling.py:163) with
    num_image_with_embedding = K and its temporal embeddings ZEROED (the reference never adds them, modeling.py:87),
    loaded with the repo's seeded encoder / projection / decoder weights, on 2 samples x 2 sampled frames."""
    from transformers import GitConfig, GitForCausalLM
    from oracle import git as git_oracle
    K, L = 2, 9
    enc_sd = synth.random_encoder_state_dict(synth.REF_SEED)
    psd = synth.random_projection_state_dict()
    dsd = synth.random_git_decoder_state_dict()
    model = GitForCausalLM(GitConfig(num_image_with_embedding=K)).eval()
    full = dict(dsd)
    full.update({"git.image_encoder." + k: v for k, v in enc_sd.items()})
    full.update({"git.visual_projection." + k: v for k, v in psd.items()})
    for f in range(K):
        full[f"git.img_temporal_embedding.{f}"] = torch.zeros(1, 1, synth.HIDDEN)
    missing, unexpected = model.load_state_dict(full, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    frames = torch.stack([vit.image_processor_224(synth.make_clip(c, K)) for c in (60, 61)])      # [2, K, 3, 224, 224]
    g = torch.Generator().manual_seed(9)
    ids = torch.randint(1000, synth.GIT_VOCAB, (2, L), generator=g)
    ids[:, 0] = 101
    mask = torch.ones(2, L, dtype=torch.long)
    mask[1, 6:] = 0                                                 # right padding of the second question
    ids[1, 6:] = 0
    with torch.no_grad():
        out = model(input_ids=ids, attention_mask=mask, pixel_values=frames)
        logits = out.logits[:, K * 197:, :]                         # text rows
        mine = git_oracle.GitVqaOracle(enc_sd, psd, dsd)(frames, ids)
    valid = mask.bool()
    print("git vqa: restatement vs HF max |d logit| on valid text rows:", float((mine - logits)[valid].abs().max()),
          "| logit std", float(logits[valid].std()))
    # the reference's loss expression (modeling.py:208-215) evaluated verbatim on HF's logits (HF 5.5's own GitForCausalLM
    # loss shifts a second time inside its loss_function, so it is not the reference's objective)
    labels = ids.clone()
    labels[~valid] = -100
    with torch.no_grad():
        shifted_logits = out.logits[:, K * 197:-1, :].contiguous()
        loss = torch.nn.CrossEntropyLoss()(shifted_logits.view(-1, synth.GIT_VOCAB), labels[:, 1:].contiguous().view(-1))
        print("git vqa: loss", float(loss), "restatement", float(git_oracle.GitVqaOracle(enc_sd, psd, dsd).loss(frames, ids, labels)))
    # greedy decoding as the reference's evaluation runs it (modeling.py:333: model.generate(**inputs, max_length=...)):
    # argmax of HF's forward, step by step WITHOUT its KV cache, on a 4-token prompt, against the restated greedy loop.
    # (HF 5.5's cached generate is not used as the pin: from the second cached step on it departs from the same model's
    # uncached forward -- checked here, it picks 487 where the forward's argmax is 5682 at a top-2 margin of 0.35 -- so it
    # is not a faithful execution of the model the reference was written against.)
    prompt = ids[:, :4].contiguous()
    with torch.no_grad():
        hf_gen = prompt.clone()
        while hf_gen.shape[1] < 12:
            step = model(input_ids=hf_gen, attention_mask=torch.ones_like(hf_gen), pixel_values=frames, use_cache=False)
            hf_gen = torch.cat([hf_gen, step.logits[:, -1].argmax(dim=-1, keepdim=True)], dim=1)
        my_gen, margins = git_oracle.GitVqaOracle(enc_sd, psd, dsd).generate(frames, prompt, max_length=12)
    print("git vqa: greedy on HF's forward", hf_gen.tolist(), "restatement equal:", torch.equal(hf_gen, my_gen),
          "min top-2 margin", float(margins.min()))
    top = logits.topk(5, dim=-1)
    np.savez_compressed(os.path.join(GOLD, "git_vqa_hf.npz"), labels=labels.numpy(), loss=np.float32(loss),
                        gen_prompt=prompt.numpy(), gen_ids=hf_gen.numpy(), gen_margins=margins.numpy(), clip_ids=np.asarray([60, 61]), K=np.int64(K),
                        input_ids=ids.numpy(), attention_mask=mask.numpy(), logits_probe=logits[:, :, ::61].numpy(),
                        logits_row=logits[0, 3].numpy(), top5_idx=top.indices.numpy(), top5_val=top.values.numpy())


if __name__ == "__main__":
    if "--resize-only" in sys.argv:
        gen_resize()
        sys.exit(0)
    if "--git-only" in sys.argv:
        gen_git_vqa()
        sys.exit(0)
    if "--scorer-only" in sys.argv:
        gen_bert_scorer()
        sys.exit(0)
    if "--visual-only" in sys.argv:
        gen_visual_tokens()
        sys.exit(0)
    assert ref_loader.available(), "needs /root/reference"
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    gen_mdf_select()
    gen_misc()
    gen_encoder_and_e2e()
    gen_resize()
    gen_visual_tokens()
    gen_bert_scorer()
    gen_git_vqa()
