"""fp32 restatement of the downstream video-QA forward on the sampled frames.  TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py).

Reference: ``src/modeling/modeling.py:29-232`` -- ``MyGitModel.forward`` (5-D ``pixel_values``: the image encoder
frame by frame, concatenation along the sequence, ``visual_projection``; NO temporal embedding, the line is
commented out at ``:87``; text embeddings appended; the combined mask of ``:120-140``) and
``MyGitForCausalLM.forward`` (``:163-232``: ``logits = self.output(sequence_output)``).  The classes it
subclasses are third-party: HF ``transformers`` ``GitModel`` (unpinned by the reference; 5.5.0 installed, where the
reference's own subclass no longer runs -- ``get_head_mask`` is gone -- so this restatement is pinned against HF's
``GitForCausalLM`` with its temporal embeddings zeroed, ``tests/golden/git_vqa_hf.npz``).  Algorithm
(``transformers/models/git/modeling_git.py``):

* ``GitEmbeddings`` (``:157-199``): word + position (0..L-1) -> LayerNorm (eps 1e-12);
* sequence = [projected visual tokens (K * 197 rows) | text rows]; mask: visual rows see the visual rows only, text
  row t sees every visual row and text rows <= t (causal), padded text keys are masked (right padding puts them
  after every valid query anyway);
* 6 post-LN blocks (``GitLayer`` = BERT layer: ``:202-388``): self-attention (scale 1/8) -> dense -> LN(x + .) ->
  dense 3072 -> erf gelu -> dense -> LN(x + .);
* output head: Linear(768, vocab) -- evaluated here on the text rows only (the reference computes the visual
  rows' logits too and never uses them: its loss slices them away, ``modeling.py:211-215``).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import vit

HIDDEN, HEADS, HEAD_DIM, EPS = 768, 12, 64, 1e-12


class GitVqaOracle:
    """``GitVqaOracle(encoder_sd, projection_sd, decoder_sd)(pixel_values [B, K, 3, 224, 224], input_ids [B, L])
    -> logits [B, L, vocab]`` (text rows)."""

    def __init__(self, encoder_sd: dict, projection_sd: dict, decoder_sd: dict, n_layers: int | None = None):
        self.enc = vit.VitOracle(encoder_sd)
        self.psd = {k: v.float() for k, v in projection_sd.items()}
        self.w = {k: v.detach().float() for k, v in decoder_sd.items()}
        self.n_layers = n_layers if n_layers is not None else 1 + max(
            int(k.split(".")[3]) for k in self.w if k.startswith("git.encoder.layer."))

    def _ln(self, x, name):
        return F.layer_norm(x, (HIDDEN,), self.w[name + ".weight"], self.w[name + ".bias"], EPS)

    def _lin(self, x, name):
        return F.linear(x, self.w[name + ".weight"], self.w[name + ".bias"])

    @torch.no_grad()
    def hidden_states(self, pixel_values: torch.Tensor, input_ids: torch.Tensor, n_layers: int | None = None):
        """Full sequence hidden state [B, K*197 + L, 768] after ``n_layers`` blocks."""
        n_layers = self.n_layers if n_layers is None else n_layers
        input_ids = torch.as_tensor(input_ids).long()
        B, L = input_ids.shape
        vis = vit.visual_tokens(pixel_values.float(), self.enc, self.psd)                 # modeling.py:76-95
        Nv = vis.shape[1]
        e = "git.embeddings."
        txt = self.w[e + "word_embeddings.weight"][input_ids] + self.w[e + "position_embeddings.weight"][:L][None]
        txt = self._ln(txt, e + "LayerNorm")
        x = torch.cat([vis, txt], dim=1)                                                  # modeling.py:114
        S = Nv + L
        q_idx = torch.arange(S)[:, None]
        k_idx = torch.arange(S)[None, :]
        allowed = torch.where(q_idx < Nv, k_idx < Nv, k_idx <= q_idx)                     # modeling.py:116-127
        bias = torch.zeros(S, S).masked_fill(~allowed, torch.finfo(torch.float32).min)
        for l in range(n_layers):
            p = f"git.encoder.layer.{l}."
            q = self._lin(x, p + "attention.self.query").view(B, S, HEADS, HEAD_DIM).transpose(1, 2)
            k = self._lin(x, p + "attention.self.key").view(B, S, HEADS, HEAD_DIM).transpose(1, 2)
            v = self._lin(x, p + "attention.self.value").view(B, S, HEADS, HEAD_DIM).transpose(1, 2)
            s = q @ k.transpose(-1, -2) / math.sqrt(HEAD_DIM) + bias
            a = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, S, HIDDEN)
            x = self._ln(x + self._lin(a, p + "attention.output.dense"), p + "attention.output.LayerNorm")
            f = F.gelu(self._lin(x, p + "intermediate.dense"))
            x = self._ln(x + self._lin(f, p + "output.dense"), p + "output.LayerNorm")
        return x, Nv

    @torch.no_grad()
    def __call__(self, pixel_values: torch.Tensor, input_ids: torch.Tensor) -> torch.Tensor:
        x, Nv = self.hidden_states(pixel_values, input_ids)
        return F.linear(x[:, Nv:], self.w["output.weight"], self.w["output.bias"])      # modeling.py:207, text rows

    @torch.no_grad()
    def loss(self, pixel_values: torch.Tensor, input_ids: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """``modeling.py:208-215``: next-token prediction -- ``shifted_logits = logits[:, num_image_tokens:-1]``,
        ``labels = labels[:, 1:]``, ``CrossEntropyLoss()`` (mean, ignore_index -100)."""
        logits = self(pixel_values, input_ids)
        shifted = logits[:, :-1, :].contiguous()
        tgt = torch.as_tensor(labels).long()[:, 1:].contiguous()
        return F.cross_entropy(shifted.view(-1, shifted.shape[-1]), tgt.view(-1))

    @torch.no_grad()
    def generate(self, pixel_values: torch.Tensor, input_ids: torch.Tensor, max_length: int = 50, eos_token_id: int = 102,
                 pad_token_id: int = 0):
        """Greedy search as HF ``generate`` runs it for ``modeling.py:333`` (``do_sample=False``, one beam): append
        ``argmax(logits[:, -1])`` until every sequence has emitted eos or ``max_length`` is reached; finished sequences
        are padded.  Also returns, per step, the top-2 logit margin of every sequence (for tie-aware comparisons)."""
        ids = torch.as_tensor(input_ids).long().clone()
        B = ids.shape[0]
        done = torch.zeros(B, dtype=torch.bool)
        margins = []
        while ids.shape[1] < max_length and not bool(done.all()):
            logits = self(pixel_values, ids)[:, -1, :]
            top2 = logits.topk(2, dim=-1).values
            margins.append((top2[:, 0] - top2[:, 1]).masked_fill(done, float("inf")))
            nxt = logits.argmax(dim=-1)
            nxt = torch.where(done, torch.full_like(nxt, pad_token_id), nxt)
            ids = torch.cat([ids, nxt[:, None]], dim=1)
            done |= nxt == eos_token_id
        return ids, (torch.stack(margins, dim=1) if margins else torch.zeros(B, 0))
