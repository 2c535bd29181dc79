"""Synthetic inputs for the frame-sampling hot path (no datasets or checkpoints offline).

Clips follow SURVEY.md section 8(d): scene-structured uint8 RGB clips laid out
``[T, H, W, 3]`` (HWC, what cv2 delivers after the BGR->RGB swap in the reference,
``src/preprocessing/prefetch_loader.py:57-67``), seeded ``666 + clip_id`` (666 is the
reference's seed, ``src/preprocessing/extract_features.py:136``).

Encoder weights are a seeded random ViT-B/16 state dict using the key names of the HF
``GitVisionModel`` the reference loads (``extract_features.py:145``), so the same dict
feeds the CUDA encoder, the oracle restatement and HF itself.
"""
from __future__ import annotations

import math

import torch

IMG = 224
PATCH = 16
HIDDEN = 768
HEADS = 12
FFN = 3072
LAYERS = 12
TOKENS = (IMG // PATCH) ** 2 + 1  # 197
REF_SEED = 666

# CLIPImageProcessor defaults (the reference's AutoProcessor)
IMAGE_MEAN = (0.48145466, 0.4578275, 0.40821073)
IMAGE_STD = (0.26862954, 0.26130258, 0.27577711)


def make_clip(clip_id: int, T: int, device="cpu", seed: int = REF_SEED, H: int = IMG, W: int = IMG,
              structured: bool = True) -> torch.Tensor:
    """One uint8 clip ``[T, H, W, 3]``.

    structured=True: ``max(2, T // 16)`` low-frequency scenes (7x7 noise, bicubic upsample),
    contiguous scene segments of random length, a slow per-frame gain drift and N(0, 0.05)
    pixel noise.  structured=False: iid uniform noise (throughput only).
    """
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed + int(clip_id))
    if T == 0:
        return torch.zeros(0, H, W, 3, dtype=torch.uint8, device=dev)
    if not structured:
        return torch.randint(0, 256, (T, H, W, 3), generator=g, device=dev, dtype=torch.uint8)
    n_scenes = max(2, T // 16)
    low = torch.rand(n_scenes, 3, 7, 7, generator=g, device=dev)
    scenes = torch.nn.functional.interpolate(low, size=(H, W), mode="bicubic", align_corners=False)
    scenes = scenes.clamp_(0.0, 1.0)
    # contiguous segments with random cut points
    w = torch.rand(n_scenes, generator=g, device=dev) + 0.25
    bounds = torch.cumsum(w / w.sum(), 0) * T
    t = torch.arange(T, device=dev, dtype=torch.float32)
    scene_of_t = torch.bucketize(t + 0.5, bounds[:-1].contiguous())
    phase = torch.rand(2, generator=g, device=dev)
    gain = 1.0 + 0.12 * torch.sin(2 * math.pi * (t / max(T, 1) * (1.0 + 2.0 * phase[0]) + phase[1]))
    out = torch.empty(T, H, W, 3, dtype=torch.uint8, device=dev)
    step = 64
    for s in range(0, T, step):
        e = min(T, s + step)
        fr = scenes[scene_of_t[s:e]] * gain[s:e, None, None, None]
        fr = fr + 0.05 * torch.randn(e - s, 3, H, W, generator=g, device=dev)
        out[s:e] = (fr.clamp_(0.0, 1.0) * 255.0).round_().to(torch.uint8).permute(0, 2, 3, 1)
    return out


def make_clips(clip_ids, T: int, device="cpu", seed: int = REF_SEED, structured: bool = True, H: int = IMG,
               W: int = IMG) -> torch.Tensor:
    """Batch of clips ``[B, T, H, W, 3]`` uint8."""
    ids = list(clip_ids)
    out = torch.empty(len(ids), T, H, W, 3, dtype=torch.uint8, device=device)
    for b, cid in enumerate(ids):
        out[b] = make_clip(cid, T, device=device, seed=seed, H=H, W=W, structured=structured)
    return out


def normalize_frames_reference(u8_hwc: torch.Tensor) -> torch.Tensor:
    """fp32 ``[T, 3, H, W]`` exactly as the HF image processor produces for 224x224 input
    (resize and centre-crop are identities there): ``(x * (1/255) - mean) / std``.
    Host/torch helper for tests and the reference arm; the product path does this in K1/K5."""
    x = u8_hwc.permute(0, 3, 1, 2).to(torch.float32) * (1.0 / 255.0)
    mean = torch.tensor(IMAGE_MEAN, dtype=torch.float32, device=x.device).view(1, 3, 1, 1)
    std = torch.tensor(IMAGE_STD, dtype=torch.float32, device=x.device).view(1, 3, 1, 1)
    return (x - mean) / std


def state_dict_keys():
    keys = [
        ("vision_model.embeddings.class_embedding", (HIDDEN,)),
        ("vision_model.embeddings.patch_embedding.weight", (HIDDEN, 3, PATCH, PATCH)),
        ("vision_model.embeddings.position_embedding.weight", (TOKENS, HIDDEN)),
        ("vision_model.pre_layrnorm.weight", (HIDDEN,)),
        ("vision_model.pre_layrnorm.bias", (HIDDEN,)),
    ]
    for l in range(LAYERS):
        p = f"vision_model.encoder.layers.{l}."
        for nm in ("k_proj", "v_proj", "q_proj", "out_proj"):
            keys.append((p + f"self_attn.{nm}.weight", (HIDDEN, HIDDEN)))
            keys.append((p + f"self_attn.{nm}.bias", (HIDDEN,)))
        keys.append((p + "layer_norm1.weight", (HIDDEN,)))
        keys.append((p + "layer_norm1.bias", (HIDDEN,)))
        keys.append((p + "mlp.fc1.weight", (FFN, HIDDEN)))
        keys.append((p + "mlp.fc1.bias", (FFN,)))
        keys.append((p + "mlp.fc2.weight", (HIDDEN, FFN)))
        keys.append((p + "mlp.fc2.bias", (HIDDEN,)))
        keys.append((p + "layer_norm2.weight", (HIDDEN,)))
        keys.append((p + "layer_norm2.bias", (HIDDEN,)))
    keys.append(("vision_model.post_layernorm.weight", (HIDDEN,)))
    keys.append(("vision_model.post_layernorm.bias", (HIDDEN,)))
    return keys


def random_encoder_state_dict(seed: int = REF_SEED, bf16_exact: bool = True) -> dict:
    """Seeded random ViT-B/16 weights under HF ``GitVisionModel`` key names (fp32, CPU).

    Unlike HF's default init, LayerNorm gains/biases and linear biases are non-trivial so
    every term of the encoder is exercised.  With ``bf16_exact`` the matrices that the CUDA
    path stores in bf16 are rounded to bf16-representable fp32 values, so the fp32 reference
    and the bf16 kernels see identical weights and differ only by activation rounding.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd = {}
    for name, shape in state_dict_keys():
        if name.endswith("norm.weight") or name.endswith("norm1.weight") or name.endswith("norm2.weight") \
                or name.endswith("layrnorm.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
        if bf16_exact and t.dim() >= 2 and "position_embedding" not in name:
            t = t.to(torch.bfloat16).to(torch.float32)
        sd[name] = t.contiguous()
    return sd


OUTLIER_CHANNELS = (37, 412, 700)


def outlier_encoder_state_dict(seed: int = REF_SEED + 7, scale: float = 50.0) -> dict:
    """Seeded ViT-B/16 weights with the activation statistics real CLIP / GIT towers have and benign random weights lack:
    a few MASSIVE residual channels (rows ``OUTLIER_CHANNELS`` of ``mlp.fc2`` in layers 2-5 scaled by ``scale``: the fp32
    residual stream then carries values two orders of magnitude above the rest through every later LayerNorm), heavy-tailed
    LayerNorm gains (log-normal, a few channels at 3-5x) and matrices that are NOT bf16-representable (the upload rounds
    them, as it would a real checkpoint).  Used by the parity stress tests only."""
    sd = random_encoder_state_dict(seed, bf16_exact=False)
    g = torch.Generator(device="cpu")
    g.manual_seed(seed + 1)
    ch = torch.tensor(OUTLIER_CHANNELS)
    for l in range(2, 6):
        p = f"vision_model.encoder.layers.{l}.mlp.fc2."
        sd[p + "weight"][ch] *= scale
        sd[p + "bias"][ch] *= scale
    for name in sd:
        if name.endswith(("layer_norm1.weight", "layer_norm2.weight", "post_layernorm.weight", "pre_layrnorm.weight")):
            sd[name] = torch.exp(0.5 * torch.randn(sd[name].shape, generator=g)).contiguous()
    return sd


def random_projection_state_dict(seed: int = REF_SEED + 1, bf16_exact: bool = True) -> dict:
    """Seeded weights of GIT's ``visual_projection`` (HF ``GitProjection``: Linear(768, 768) + LayerNorm) under its
    own key names; the matrix is bf16-representable like the encoder matrices."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    w = 0.03 * torch.randn(HIDDEN, HIDDEN, generator=g)
    if bf16_exact:
        w = w.to(torch.bfloat16).to(torch.float32)
    return {
        "visual_projection.0.weight": w.contiguous(),
        "visual_projection.0.bias": 0.02 * torch.randn(HIDDEN, generator=g),
        "visual_projection.1.weight": 1.0 + 0.1 * torch.randn(HIDDEN, generator=g),
        "visual_projection.1.bias": 0.02 * torch.randn(HIDDEN, generator=g),
    }


def question_embeddings(clip_ids, seed: int = REF_SEED, device="cpu") -> torch.Tensor:
    """Synthetic question text embeddings for the MIF workload (BASELINE config 3): one unit vector
    ``normalize(randn(768))`` per clip from ``Generator(seed + clip_id)`` -- the embedding-space surrogate of
    SURVEY.md 8(d); the reference's own relevance model (a caption cross-encoder) is out of scope."""
    rows = []
    for cid in clip_ids:
        g = torch.Generator()
        g.manual_seed(seed + int(cid))
        rows.append(torch.nn.functional.normalize(torch.randn(HIDDEN, generator=g), dim=0))
    return torch.stack(rows).to(device) if rows else torch.zeros(0, HIDDEN, device=device)


# ---------------------------------------------------------------------------------------------
# MIF relevance model (gen_sample.py:113,160: a bert-base-cased sequence classifier) -- synthetic
# weights, tokenizer and (question, captions) workload; no checkpoint or vocabulary offline.
# ---------------------------------------------------------------------------------------------
BERT_VOCAB = 28996        # bert-base-cased
BERT_MAX_POS = 512
BERT_TYPES = 2
BERT_LABELS = 2
BERT_CLS, BERT_SEP, BERT_PAD = 101, 102, 0


def scorer_state_dict_keys(vocab: int = BERT_VOCAB, labels: int = BERT_LABELS):
    """HF ``BertForSequenceClassification.state_dict()`` key order (floating-point entries)."""
    e = "bert.embeddings."
    keys = [
        (e + "word_embeddings.weight", (vocab, HIDDEN)),
        (e + "position_embeddings.weight", (BERT_MAX_POS, HIDDEN)),
        (e + "token_type_embeddings.weight", (BERT_TYPES, HIDDEN)),
        (e + "LayerNorm.weight", (HIDDEN,)),
        (e + "LayerNorm.bias", (HIDDEN,)),
    ]
    for l in range(LAYERS):
        p = f"bert.encoder.layer.{l}."
        for nm, shape in (("attention.self.query", (HIDDEN, HIDDEN)), ("attention.self.key", (HIDDEN, HIDDEN)),
                          ("attention.self.value", (HIDDEN, HIDDEN)), ("attention.output.dense", (HIDDEN, HIDDEN))):
            keys.append((p + nm + ".weight", shape))
            keys.append((p + nm + ".bias", (shape[0],)))
        keys.append((p + "attention.output.LayerNorm.weight", (HIDDEN,)))
        keys.append((p + "attention.output.LayerNorm.bias", (HIDDEN,)))
        keys.append((p + "intermediate.dense.weight", (FFN, HIDDEN)))
        keys.append((p + "intermediate.dense.bias", (FFN,)))
        keys.append((p + "output.dense.weight", (HIDDEN, FFN)))
        keys.append((p + "output.dense.bias", (HIDDEN,)))
        keys.append((p + "output.LayerNorm.weight", (HIDDEN,)))
        keys.append((p + "output.LayerNorm.bias", (HIDDEN,)))
    keys += [("bert.pooler.dense.weight", (HIDDEN, HIDDEN)), ("bert.pooler.dense.bias", (HIDDEN,)),
             ("classifier.weight", (labels, HIDDEN)), ("classifier.bias", (labels,))]
    return keys


def random_scorer_state_dict(seed: int = REF_SEED + 2, vocab: int = BERT_VOCAB, labels: int = BERT_LABELS,
                             bf16_exact: bool = True) -> dict:
    """Seeded random bert-base sequence-classifier weights under HF key names (fp32, CPU).  Non-trivial LayerNorm
    gains/biases and linear biases; the encoder-layer matrices (stored in bf16 on the GPU) are bf16-representable
    with ``bf16_exact`` so the fp32 oracle and the kernels see identical weights."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd = {}
    for name, shape in scorer_state_dict_keys(vocab, labels):
        if name.endswith("LayerNorm.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        elif name.startswith("classifier") or "pooler" in name:
            t = 0.05 * torch.randn(shape, generator=g)
        elif "embeddings" in name:
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            t = 0.04 * torch.randn(shape, generator=g)
            if bf16_exact:
                t = t.to(torch.bfloat16).to(torch.float32)
        sd[name] = t.contiguous()
    return sd


class SynthTokenizer:
    """Stand-in for ``AutoTokenizer.from_pretrained(args.sim_model)`` (gen_sample.py:159) with the call signature
    the reference uses (gen_sample.py:80): whitespace words hashed into the vocabulary, ``[CLS] a [SEP] b [SEP]``,
    token types 0 / 1, right padding to the longest pair of the call, ``truncation`` to ``max_length`` by trimming
    the longer side first (HF 'longest_first').  Returns int64 tensors like ``return_tensors='pt'``."""

    def __init__(self, vocab: int = BERT_VOCAB, max_length: int = BERT_MAX_POS):
        self.vocab, self.max_length = vocab, max_length

    def _ids(self, text: str):
        import zlib
        lo = min(1000, self.vocab // 2)
        return [lo + zlib.crc32(w.encode()) % (self.vocab - lo) for w in text.split()]

    def __call__(self, text, text_pair=None, padding=True, truncation=True, return_tensors="pt", max_length=None):
        max_length = max_length or self.max_length
        rows = []
        for i, a in enumerate(text):
            ta = self._ids(a)
            tb = self._ids(text_pair[i]) if text_pair is not None else None
            if truncation:
                budget = max_length - (3 if tb is not None else 2)
                while len(ta) + (len(tb) if tb else 0) > budget:
                    if tb and len(tb) >= len(ta):
                        tb.pop()
                    else:
                        ta.pop()
            ids = [BERT_CLS] + ta + [BERT_SEP]
            types = [0] * len(ids)
            if tb is not None:
                ids += tb + [BERT_SEP]
                types += [1] * (len(tb) + 1)
            rows.append((ids, types))
        L = max(len(r[0]) for r in rows) if rows else 0
        input_ids = torch.full((len(rows), L), BERT_PAD, dtype=torch.long)
        token_type_ids = torch.zeros(len(rows), L, dtype=torch.long)
        attention_mask = torch.zeros(len(rows), L, dtype=torch.long)
        for i, (ids, types) in enumerate(rows):
            input_ids[i, :len(ids)] = torch.tensor(ids)
            token_type_ids[i, :len(ids)] = torch.tensor(types)
            attention_mask[i, :len(ids)] = 1
        return {"input_ids": input_ids, "token_type_ids": token_type_ids, "attention_mask": attention_mask}


_WORDS = ("a man woman dog cat child person is are playing cutting riding running cooking holding the guitar piano "
          "horse bike ball food knife water street kitchen field room with on in near two people someone what who how "
          "many does doing where red blue small large quickly slowly table car road grass").split()


def make_qa_workload(n_samples: int, n_captions: int, seed: int = REF_SEED, min_words: int = 4, max_words: int = 14):
    """Synthetic MSVD-QA-shaped annotations for the MIF step: ``(qa_samples, all_captions)`` --
    ``qa_samples[i] = {'video': i, 'question': str, 'answer': str}`` and ``all_captions['video{i}']`` = one caption
    per sampled frame (what gen_sample.py:20-45 writes to frame_captions.json)."""
    g = torch.Generator()
    g.manual_seed(seed)

    def sentence(lo, hi):
        n = int(torch.randint(lo, hi + 1, (1,), generator=g))
        return " ".join(_WORDS[int(j)] for j in torch.randint(0, len(_WORDS), (n,), generator=g))

    qa, caps = [], {}
    for i in range(n_samples):
        qa.append({"video": i, "question": sentence(4, 10) + " ?", "answer": _WORDS[i % len(_WORDS)]})
        caps[f"video{i}"] = [sentence(min_words, max_words) for _ in range(n_captions)]
    return qa, caps


# ---------------------------------------------------------------------------------------------
# Downstream video-QA model (src/modeling/modeling.py: MyGitForCausalLM, git-base geometry): the text
# side -- embeddings, 6 post-LN blocks over [visual tokens | text], output head -- with seeded weights.
# ---------------------------------------------------------------------------------------------
GIT_VOCAB = 30522
GIT_MAX_POS = 1024
GIT_LAYERS = 6


def git_decoder_state_dict_keys(vocab: int = GIT_VOCAB, n_layers: int = GIT_LAYERS, max_pos: int = GIT_MAX_POS):
    """Decoder-side entries of HF ``GitForCausalLM.state_dict()`` in its own order (the image encoder and the visual
    projection sit between the embeddings and the layers there; they are loaded into the FrameEncoder)."""
    keys = [
        ("git.embeddings.word_embeddings.weight", (vocab, HIDDEN)),
        ("git.embeddings.position_embeddings.weight", (max_pos, HIDDEN)),
        ("git.embeddings.LayerNorm.weight", (HIDDEN,)),
        ("git.embeddings.LayerNorm.bias", (HIDDEN,)),
    ]
    for l in range(n_layers):
        p = f"git.encoder.layer.{l}."
        for nm, shape in (("attention.self.query", (HIDDEN, HIDDEN)), ("attention.self.key", (HIDDEN, HIDDEN)),
                          ("attention.self.value", (HIDDEN, HIDDEN)), ("attention.output.dense", (HIDDEN, HIDDEN))):
            keys.append((p + nm + ".weight", shape))
            keys.append((p + nm + ".bias", (shape[0],)))
        keys.append((p + "attention.output.LayerNorm.weight", (HIDDEN,)))
        keys.append((p + "attention.output.LayerNorm.bias", (HIDDEN,)))
        keys.append((p + "intermediate.dense.weight", (FFN, HIDDEN)))
        keys.append((p + "intermediate.dense.bias", (FFN,)))
        keys.append((p + "output.dense.weight", (HIDDEN, FFN)))
        keys.append((p + "output.dense.bias", (HIDDEN,)))
        keys.append((p + "output.LayerNorm.weight", (HIDDEN,)))
        keys.append((p + "output.LayerNorm.bias", (HIDDEN,)))
    keys += [("output.weight", (vocab, HIDDEN)), ("output.bias", (vocab,))]
    return keys


def random_git_decoder_state_dict(seed: int = REF_SEED + 3, vocab: int = GIT_VOCAB, n_layers: int = GIT_LAYERS,
                                  bf16_exact: bool = True) -> dict:
    """Seeded random weights of the GIT text decoder + output head under HF key names (fp32, CPU); the matrices the
    GPU keeps in bf16 (layer projections, MLP, output head) are bf16-representable with ``bf16_exact``."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd = {}
    for name, shape in git_decoder_state_dict_keys(vocab, n_layers):
        if name.endswith("LayerNorm.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        elif "embeddings" in name:
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            t = 0.04 * torch.randn(shape, generator=g)
            if bf16_exact:
                t = t.to(torch.bfloat16).to(torch.float32)
        sd[name] = t.contiguous()
    return sd
