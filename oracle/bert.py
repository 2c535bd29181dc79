"""fp32 restatement of the MIF relevance model (a BERT sequence classifier) and of the loop body
that calls it.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference call site, ``src/preprocessing/gen_sample.py:79-88``::

    inputs = tokenizer(text=[question]*bsz, text_pair=captions, padding=True, truncation=True, return_tensors='pt')
    output = model(**inputs)                # AutoModelForSequenceClassification (:160)
    scores = output[0][:,0]                 # logits of label 0
    inds = scores[::ds_rate].topk(args.K)[1]; inds = [i*ds_rate for i in inds]

The model is third-party: HF ``transformers`` (unpinned by the reference; 5.5.0 installed)
``BertForSequenceClassification`` for ``iarfmoose/bert-base-cased-qa-evaluator``
(``gen_sample.py:113``) -- bert-base-cased geometry: vocab 28996, 768 hidden, 12 layers x 12
heads, FFN 3072, 512 positions, 2 token types, LayerNorm eps 1e-12, exact (erf) GELU, 2 labels.
The checkpoint is not reachable offline, so parity is pinned on seeded random weights only
("parity unpinned" for the real checkpoint).  Algorithm as in
``transformers/models/bert/modeling_bert.py``:

* embeddings (``:53-140``): word + position (0..L-1) + token-type rows, LayerNorm;
* 12 post-LN blocks (``:143-420``): self-attention (scale 1/8, additive key mask, fp32 softmax)
  -> dense -> LayerNorm(x + .) -> dense 3072 -> gelu -> dense -> LayerNorm(x + .);
* pooler (``:456-468``): tanh(dense(hidden[:, 0])); classifier (``:1077-1155``): Linear(768, labels).

``tests/test_oracle_golden.py`` checks this against HF itself (live when importable, and through
``tests/golden/bert_scorer_hf.npz``).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

HIDDEN, HEADS, HEAD_DIM, LAYERS, FFN, EPS = 768, 12, 64, 12, 3072, 1e-12


class BertScorerOracle:
    """``BertScorerOracle(sd)(input_ids, token_type_ids, attention_mask) -> logits [N, labels]``."""

    def __init__(self, state_dict: dict, dtype=torch.float32):
        self.w = {k: v.detach().to(dtype) for k, v in state_dict.items() if v.is_floating_point()}
        self.dtype = dtype

    def _ln(self, x, name):
        return F.layer_norm(x, (HIDDEN,), self.w[name + ".weight"], self.w[name + ".bias"], EPS)

    def _lin(self, x, name):
        return F.linear(x, self.w[name + ".weight"], self.w[name + ".bias"])

    def hidden_states(self, input_ids, token_type_ids=None, attention_mask=None, n_layers: int = LAYERS):
        input_ids = torch.as_tensor(input_ids).long()
        N, L = input_ids.shape
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_ids)
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        token_type_ids = torch.as_tensor(token_type_ids).long()
        attention_mask = torch.as_tensor(attention_mask)
        e = "bert.embeddings."
        x = self.w[e + "word_embeddings.weight"][input_ids] + self.w[e + "position_embeddings.weight"][:L][None] \
            + self.w[e + "token_type_embeddings.weight"][token_type_ids]
        x = self._ln(x, e + "LayerNorm")
        # additive key mask, broadcast over heads and queries (masked keys get the most negative value)
        bias = (1.0 - attention_mask.to(self.dtype))[:, None, None, :] * torch.finfo(self.dtype).min
        for l in range(n_layers):
            p = f"bert.encoder.layer.{l}."
            q = self._lin(x, p + "attention.self.query").view(N, L, HEADS, HEAD_DIM).transpose(1, 2)
            k = self._lin(x, p + "attention.self.key").view(N, L, HEADS, HEAD_DIM).transpose(1, 2)
            v = self._lin(x, p + "attention.self.value").view(N, L, HEADS, HEAD_DIM).transpose(1, 2)
            s = q @ k.transpose(-1, -2) / math.sqrt(HEAD_DIM) + bias
            a = torch.softmax(s, dim=-1) @ v
            a = a.transpose(1, 2).reshape(N, L, HIDDEN)
            x = self._ln(x + self._lin(a, p + "attention.output.dense"), p + "attention.output.LayerNorm")
            f = F.gelu(self._lin(x, p + "intermediate.dense"))
            x = self._ln(x + self._lin(f, p + "output.dense"), p + "output.LayerNorm")
        return x

    def __call__(self, input_ids, token_type_ids=None, attention_mask=None):
        x = self.hidden_states(input_ids, token_type_ids, attention_mask)
        pooled = torch.tanh(self._lin(x[:, 0], "bert.pooler.dense"))
        return self._lin(pooled, "classifier")


def mif_indices_from_logits(logits: torch.Tensor, K: int, ds_rate: int = 1, label: int = 0):
    """``gen_sample.py:83-88``: scores = logits[:, 0]; strided top-K, indices scaled back, best first."""
    scores = logits[:, label]
    inds = scores[::ds_rate].topk(K)[1].tolist()
    return [i * ds_rate for i in inds]


def generate_inds(tokenizer, model, qa_samples, all_captions, K: int, ds_rate: int = 1, vid_name: str = "video",
                  qid_temp: str = "video{}"):
    """Loop body of ``generate_inds`` (``gen_sample.py:68-91``), one QA sample at a time like the reference:
    ``model`` maps the tokenizer's dict to logits.  Returns the new list with ``sampled_inds``."""
    out = []
    for sample in qa_samples:
        captions = all_captions[qid_temp.format(sample[vid_name])]
        inputs = tokenizer(text=[sample["question"]] * len(captions), text_pair=captions, padding=True,
                           truncation=True, return_tensors="pt")
        logits = model(inputs["input_ids"], inputs.get("token_type_ids"), inputs["attention_mask"])
        new = dict(sample)
        new["sampled_inds"] = mif_indices_from_logits(logits, K, ds_rate)
        out.append(new)
    return out
