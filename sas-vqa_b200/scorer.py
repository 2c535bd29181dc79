"""Host-side mirror of the reference's MIF step (``src/preprocessing/gen_sample.py:48-94``): the caption
cross-encoder that scores (question, caption) pairs and the strided top-K that follows, on the CUDA library.

``CaptionScorer`` stands where the reference has ``AutoModelForSequenceClassification.from_pretrained(
args.sim_model).eval().cuda()`` (gen_sample.py:160,49); ``generate_inds`` keeps the reference function's name,
argument meaning and output record (``sample['sampled_inds']``, best first, gen_sample.py:86-92).  There is no
CPU fallback: everything below raises if libsasvqa_b200.so is missing.
"""
from __future__ import annotations

import ctypes
from types import SimpleNamespace

import torch

from . import _capi
from .synth import BERT_LABELS, BERT_MAX_POS, HIDDEN, scorer_state_dict_keys

PROFILE_KINDS = ("embed", "layernorm", "gemm_qkv", "attention", "gemm_out_proj", "gemm_fc1", "gemm_fc2", "pooler")


def flatten_scorer_state_dict(state_dict: dict):
    """fp32 CPU vector in the key order sasvqa_scorer_create expects; returns (flat, vocab, labels)."""
    try:
        vocab = int(state_dict["bert.embeddings.word_embeddings.weight"].shape[0])
        labels = int(state_dict["classifier.weight"].shape[0])
    except KeyError as exc:
        raise KeyError(f"scorer state dict lacks {exc.args[0]!r}; expected a BertForSequenceClassification state dict")
    parts = []
    for name, shape in scorer_state_dict_keys(vocab, labels):
        if name not in state_dict:
            raise KeyError(f"scorer state dict lacks {name!r}; expected a BertForSequenceClassification state dict")
        t = state_dict[name].detach().to("cpu", torch.float32)
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name}: shape {tuple(t.shape)} != {tuple(shape)} (only bert-base geometry is supported)")
        parts.append(t.reshape(-1))
    return torch.cat(parts).contiguous(), vocab, labels


class CaptionScorer:
    """Owns a SasvqaScorer handle (bf16 layer weights, fp32 embeddings / pooler / classifier, workspace, TMA
    descriptors on one GPU).  Callable like the reference's model: ``scorer(**tokenizer_output)`` returns an object
    whose ``[0]`` / ``.logits`` is the ``[N, num_labels]`` fp32 logits tensor on the GPU (gen_sample.py:82-83)."""

    def __init__(self, state_dict: dict, max_tokens: int = 0, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        flat, self.vocab, self.labels = flatten_scorer_state_dict(state_dict)
        lib = _capi.lib()
        assert lib.sasvqa_scorer_num_params(self.vocab, self.labels) == flat.numel()
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            rc = lib.sasvqa_scorer_create(flat.data_ptr(), flat.numel(), self.vocab, self.labels, int(max_tokens),
                                          ctypes.byref(handle))
        _capi.check(rc, "sasvqa_scorer_create")
        self._h = handle
        self.max_tokens = lib.sasvqa_scorer_max_tokens(self._h)

    @classmethod
    def from_model(cls, model, max_tokens: int = 0, device=None) -> "CaptionScorer":
        """From an HF BertForSequenceClassification (optionally wrapped in DataParallel)."""
        return cls(getattr(model, "module", model).state_dict(), max_tokens=max_tokens, device=device)

    @property
    def handle(self):
        if self._h is None:
            raise _capi.SasvqaError("scorer handle already closed")
        return self._h

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            _capi.lib().sasvqa_scorer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # the reference calls model.eval().cuda() on the object it loads (gen_sample.py:49)
    def eval(self):
        return self

    def cuda(self):
        return self

    # ---- device entry ----------------------------------------------------------------------
    def _lengths(self, input_ids, attention_mask):
        N, L = input_ids.shape
        if attention_mask is None:
            return torch.full((N,), L, dtype=torch.int32)
        m = attention_mask.detach().to("cpu").ne(0)
        lens = m.sum(dim=1).to(torch.int32)
        prefix = torch.arange(L)[None, :] < lens[:, None]
        if not torch.equal(m, prefix):
            raise ValueError("attention_mask must be right-padded (ones then zeros), as the BERT tokenizer pads")
        return lens

    def logits(self, input_ids, token_type_ids=None, attention_mask=None) -> torch.Tensor:
        """[N, L] ids (any integer dtype, CPU or GPU) -> fp32 logits [N, num_labels] on the GPU."""
        if input_ids.dim() != 2:
            raise ValueError(f"input_ids must be [N, L], got {tuple(input_ids.shape)}")
        N, L = (int(v) for v in input_ids.shape)
        if L > BERT_MAX_POS:
            raise ValueError(f"sequence length {L} exceeds the {BERT_MAX_POS}-entry position table")
        lens = self._lengths(input_ids, attention_mask)
        ids = input_ids.to(device=self.device, dtype=torch.int32).contiguous()
        tts = None if token_type_ids is None else token_type_ids.to(device=self.device, dtype=torch.int32).contiguous()
        out = torch.empty(N, self.labels, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib().sasvqa_scorer_logits(self.handle, ids.data_ptr(), _capi.ptr(tts), lens.data_ptr(), N, L,
                                                         out.data_ptr(), torch.cuda.current_stream().cuda_stream),
                        "sasvqa_scorer_logits")
        return out

    def hidden(self, input_ids, token_type_ids=None, attention_mask=None, n_layers: int = 12) -> torch.Tensor:
        """Inspection: packed fp32 hidden state [sum(lengths), 768] after ``n_layers`` blocks."""
        N, L = (int(v) for v in input_ids.shape)
        lens = self._lengths(input_ids, attention_mask)
        ids = input_ids.to(device=self.device, dtype=torch.int32).contiguous()
        tts = None if token_type_ids is None else token_type_ids.to(device=self.device, dtype=torch.int32).contiguous()
        out = torch.empty(int(lens.sum()), HIDDEN, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib().sasvqa_scorer_hidden(self.handle, ids.data_ptr(), _capi.ptr(tts), lens.data_ptr(), N, L,
                                                         int(n_layers), out.data_ptr(),
                                                         torch.cuda.current_stream().cuda_stream), "sasvqa_scorer_hidden")
        return out

    def __call__(self, input_ids=None, token_type_ids=None, attention_mask=None, **_unused):
        lg = self.logits(input_ids, token_type_ids, attention_mask)
        out = SimpleNamespace(logits=lg)
        return _Indexable(out)

    # ---- host entries ------------------------------------------------------------------------
    @staticmethod
    def _host_i64(t, name):
        if t is None:
            return None
        if t.is_cuda:
            raise ValueError(f"{name}: the host entry takes the tokenizer's CPU tensors")
        return t.to(torch.int64).contiguous()

    def logits_host(self, input_ids, token_type_ids=None, attention_mask=None) -> torch.Tensor:
        """The tokenizer's CPU int64 tensors as they are -> fp32 logits [N, num_labels] on the CPU."""
        ids = self._host_i64(input_ids, "input_ids")
        tts = self._host_i64(token_type_ids, "token_type_ids")
        msk = self._host_i64(attention_mask, "attention_mask")
        N, L = (int(v) for v in ids.shape)
        out = torch.empty(N, self.labels, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib().sasvqa_scorer_logits_host(self.handle, ids.data_ptr(), _capi.ptr(tts), _capi.ptr(msk), N, L,
                                                              out.data_ptr()), "sasvqa_scorer_logits_host")
        return out

    def select_captions_host(self, input_ids, token_type_ids, attention_mask, n_samples: int, K: int, ds_rate: int = 1,
                             label: int = 0, want_scores: bool = False, idx_out=None):
        """G = n_samples QA samples of T captions each, rows ``g*T + t`` of the [G*T, L] tokenizer arrays:
        returns (idx int32 [G, K] best first, scores [G, T] or None) on the CPU (gen_sample.py:83-88)."""
        ids = self._host_i64(input_ids, "input_ids")
        tts = self._host_i64(token_type_ids, "token_type_ids")
        msk = self._host_i64(attention_mask, "attention_mask")
        N, L = (int(v) for v in ids.shape)
        G = int(n_samples)
        if G <= 0 or N % G:
            raise ValueError(f"{N} sequences do not split into {G} samples of equal caption count")
        T = N // G
        idx = torch.empty(G, K, dtype=torch.int32) if idx_out is None else idx_out
        scores = torch.empty(G, T, dtype=torch.float32) if want_scores else None
        with torch.cuda.device(self.device):
            rc = _capi.lib().sasvqa_mif_select_captions_host(self.handle, ids.data_ptr(), _capi.ptr(tts), _capi.ptr(msk), G, T, L,
                                                             int(K), int(ds_rate), int(label), idx.data_ptr(),
                                                             _capi.ptr(scores))
        if rc == 1 and K > (T + ds_rate - 1) // ds_rate:
            raise RuntimeError("selected index k out of range")          # what torch.topk raises in the reference
        _capi.check(rc, "sasvqa_mif_select_captions_host")
        return idx, scores

    # ---- instrumentation -----------------------------------------------------------------------
    def profile(self, on: bool) -> None:
        _capi.check(_capi.lib().sasvqa_scorer_profile_enable(self.handle, int(on)), "sasvqa_scorer_profile_enable")

    def profile_read(self) -> dict:
        n = len(PROFILE_KINDS)
        ms = (ctypes.c_double * n)()
        cnt = (ctypes.c_int64 * n)()
        _capi.check(_capi.lib().sasvqa_scorer_profile_read(self.handle, ms, cnt, n), "sasvqa_scorer_profile_read")
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(PROFILE_KINDS)}


class _Indexable:
    """``output[0]`` and ``output.logits`` like an HF SequenceClassifierOutput."""

    def __init__(self, ns):
        self.logits = ns.logits

    def __getitem__(self, i):
        return (self.logits,)[i]


def move_to_cuda(inputs: dict) -> dict:
    """gen_sample.py:47-48."""
    return {k: v.cuda() for k, v in inputs.items()}


def generate_inds(tokenizer, model, qa_samples: list, all_captions: dict, K: int, ds_rate: int = 1, dataset: str = "msvd_qa",
                  samples_per_call: int = 64) -> list:
    """The loop of ``generate_inds`` (gen_sample.py:50-92) over one split's QA list: every sample's question is
    paired with the captions of its video's sampled frames, the pairs are scored by ``model`` (a ``CaptionScorer``)
    and the K best of every ``ds_rate``-th caption are stored best first under ``sampled_inds``.

    The reference runs one tokenizer call and one model call per QA sample; here ``samples_per_call`` samples with
    the same caption count are tokenized together and go through ONE library call (cross-encoder + top-K on the
    GPU).  Padding to the longest pair of the larger batch does not change a logit: padded positions are dropped
    before the first kernel.  Returns the new list (what the reference saves as ``qa_winds_{split}.json``)."""
    if dataset == "msvd_qa":
        vid_name, qid_temp = "video", "video{}"
    elif dataset == "msrvtt_qa":
        vid_name, qid_temp = "video_id", "{}"
    else:
        raise ValueError("Invalid dataset name! Current supported dataset msvd_qa, msrvtt_qa")
    if not hasattr(model, "select_captions_host"):          # an HF BertForSequenceClassification: upload it once
        model = CaptionScorer.from_model(model)
    new_ds = [None] * len(qa_samples)
    # group consecutive samples with equal caption counts (one H5 => one K for every video; be general anyway)
    order = sorted(range(len(qa_samples)), key=lambda i: len(all_captions[qid_temp.format(qa_samples[i][vid_name])]))
    pos = 0
    while pos < len(order):
        T = len(all_captions[qid_temp.format(qa_samples[order[pos]][vid_name])])
        group = []
        while pos < len(order) and len(group) < samples_per_call and \
                len(all_captions[qid_temp.format(qa_samples[order[pos]][vid_name])]) == T:
            group.append(order[pos])
            pos += 1
        text, pair = [], []
        for i in group:
            sample = qa_samples[i]
            captions = all_captions[qid_temp.format(sample[vid_name])]
            text += [sample["question"]] * len(captions)
            pair += list(captions)
        inputs = tokenizer(text=text, text_pair=pair, padding=True, truncation=True, return_tensors="pt")
        idx, _ = model.select_captions_host(inputs["input_ids"], inputs.get("token_type_ids"), inputs["attention_mask"],
                                            len(group), K, ds_rate)
        for row, i in enumerate(group):
            rec = dict(qa_samples[i])
            rec["sampled_inds"] = [int(v) for v in idx[row]]
            new_ds[i] = rec
    return new_ds
