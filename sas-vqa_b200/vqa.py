"""Host-side mirror of the reference's downstream video-QA forward on the sampled frames
(``src/modeling/modeling.py``: ``MyGitModel`` / ``MyGitForCausalLM``, wrapped by ``GITBaseModel`` at ``:314-332``):
``pixel_values`` [B, K, 3, 224, 224] (what the collator builds from the ``sampled_frames`` rows) and ``input_ids``
[B, L] in, next-token logits of the text positions out.  The visual side runs on the sampler's ``FrameEncoder``
(encoder + ``visual_projection``), the text side on a ``GitDecoder`` handle; no CPU fallback.
"""
from __future__ import annotations

import ctypes

import torch

from . import _capi
from .ops import FrameEncoder
from .synth import HIDDEN, IMG, TOKENS, git_decoder_state_dict_keys


def flatten_git_decoder_state_dict(state_dict: dict):
    """fp32 CPU vector in the key order sasvqa_git_decoder_create expects; returns (flat, vocab, n_layers)."""
    try:
        vocab = int(state_dict["output.weight"].shape[0])
    except KeyError as exc:
        raise KeyError(f"decoder state dict lacks {exc.args[0]!r}; expected a GitForCausalLM state dict")
    n_layers = 1 + max(int(k.split(".")[3]) for k in state_dict if k.startswith("git.encoder.layer."))
    parts = []
    for name, shape in git_decoder_state_dict_keys(vocab, n_layers):
        if name not in state_dict:
            raise KeyError(f"decoder state dict lacks {name!r}; expected a GitForCausalLM state dict")
        t = state_dict[name].detach().to("cpu", torch.float32)
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name}: shape {tuple(t.shape)} != {tuple(shape)} (only git-base geometry is supported)")
        parts.append(t.reshape(-1))
    return torch.cat(parts).contiguous(), vocab, n_layers


class GitDecoder:
    """Text side of the GIT video-QA model on one GPU: bf16 layer / head weights, fp32 embeddings, workspace."""

    def __init__(self, state_dict: dict, max_rows: int = 0, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        flat, self.vocab, self.n_layers = flatten_git_decoder_state_dict(state_dict)
        lib = _capi.lib()
        assert lib.sasvqa_git_decoder_num_params(self.vocab, self.n_layers) == flat.numel()
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            rc = lib.sasvqa_git_decoder_create(flat.data_ptr(), flat.numel(), self.vocab, self.n_layers, int(max_rows),
                                               ctypes.byref(handle))
        _capi.check(rc, "sasvqa_git_decoder_create")
        self._h = handle
        self.vocab_padded = lib.sasvqa_git_decoder_vocab_padded(self._h)

    @property
    def handle(self):
        if self._h is None:
            raise _capi.SasvqaError("decoder handle already closed")
        return self._h

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            _capi.lib().sasvqa_git_decoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _prep(enc: FrameEncoder, pixel_values: torch.Tensor, input_ids: torch.Tensor):
    if pixel_values.dim() != 5 or tuple(pixel_values.shape[2:]) != (3, IMG, IMG):
        raise ValueError("pixel_values must be of rank 5: (batch_size, num_frames, 3, 224, 224)")
    if input_ids.dim() != 2 or input_ids.shape[0] != pixel_values.shape[0]:
        raise ValueError(f"input_ids must be [B, L] with B = {pixel_values.shape[0]}, got {tuple(input_ids.shape)}")
    px = pixel_values.to(device=enc.device, dtype=torch.float32).contiguous()
    ids = input_ids.to(device=enc.device, dtype=torch.int32).contiguous()
    return px, ids


def vqa_logits(pixel_values: torch.Tensor, input_ids: torch.Tensor, enc: FrameEncoder, dec: GitDecoder) -> torch.Tensor:
    """``MyGitForCausalLM(input_ids=..., pixel_values=...).logits[:, K*197:, :]`` (modeling.py:163-232): fp32
    [B, L, vocab] on the GPU.  ``enc`` must carry the model's ``visual_projection`` (``FrameEncoder.set_projection``);
    padded text positions give rows the caller ignores, exactly as in the reference."""
    px, ids = _prep(enc, pixel_values, input_ids)
    B, K = int(px.shape[0]), int(px.shape[1])
    L = int(ids.shape[1])
    out = torch.empty(B, L, dec.vocab_padded, dtype=torch.float32, device=enc.device)
    with torch.cuda.device(enc.device):
        _capi.check(_capi.lib().sasvqa_git_vqa_logits_f32(dec.handle, enc.handle, px.data_ptr(), B, K, ids.data_ptr(), L,
                                                          out.data_ptr(), torch.cuda.current_stream().cuda_stream),
                    "sasvqa_git_vqa_logits_f32")
    return out[:, :, :dec.vocab]


def vqa_loss(pixel_values: torch.Tensor, input_ids: torch.Tensor, labels: torch.Tensor, enc: FrameEncoder, dec: GitDecoder,
             want_logits: bool = False):
    """``MyGitForCausalLM(input_ids=..., pixel_values=..., labels=...).loss`` (modeling.py:208-215): the mean next-token
    cross-entropy over the text rows, labels of -100 ignored; a 0-d fp32 tensor on the GPU (plus the logits on request)."""
    px, ids = _prep(enc, pixel_values, input_ids)
    if tuple(labels.shape) != tuple(input_ids.shape):
        raise ValueError(f"labels must have the shape of input_ids {tuple(input_ids.shape)}, got {tuple(labels.shape)}")
    lab = labels.to(device=enc.device, dtype=torch.int32).contiguous()
    B, K = int(px.shape[0]), int(px.shape[1])
    L = int(ids.shape[1])
    loss = torch.empty(1, dtype=torch.float32, device=enc.device)
    logits = torch.empty(B, L, dec.vocab_padded, dtype=torch.float32, device=enc.device) if want_logits else None
    with torch.cuda.device(enc.device):
        _capi.check(_capi.lib().sasvqa_git_vqa_loss_f32(dec.handle, enc.handle, px.data_ptr(), B, K, ids.data_ptr(), lab.data_ptr(),
                                                        L, loss.data_ptr(), _capi.ptr(logits),
                                                        torch.cuda.current_stream().cuda_stream), "sasvqa_git_vqa_loss_f32")
    return (loss[0], logits[:, :, :dec.vocab]) if want_logits else loss[0]


def vqa_generate(pixel_values: torch.Tensor, input_ids: torch.Tensor, enc: FrameEncoder, dec: GitDecoder, max_length: int = 50,
                 eos_token_id: int = 102, pad_token_id: int = 0, trim: bool = True) -> torch.Tensor:
    """``MyGitForCausalLM.generate(pixel_values=..., input_ids=..., max_length=50)`` as the reference's evaluation calls it
    (modeling.py:330-333): greedy search from the prompt ``input_ids`` [B, L0] (equal lengths).  Returns int64 ids
    [B, <= max_length] on the GPU: prompt, generated tokens, pad after a sequence's eos; with ``trim`` the columns after
    the step at which every sequence had finished are dropped, as HF stops there."""
    px, ids = _prep(enc, pixel_values, input_ids)
    B, K = int(px.shape[0]), int(px.shape[1])
    L0 = int(ids.shape[1])
    if not 1 <= L0 <= max_length:
        raise ValueError(f"need 1 <= prompt length ({L0}) <= max_length ({max_length})")
    out = torch.empty(B, max_length, dtype=torch.int32, device=enc.device)
    with torch.cuda.device(enc.device):
        _capi.check(_capi.lib().sasvqa_git_vqa_generate_f32(dec.handle, enc.handle, px.data_ptr(), B, K, ids.data_ptr(), L0,
                                                            int(max_length), int(eos_token_id), int(pad_token_id), out.data_ptr(),
                                                            torch.cuda.current_stream().cuda_stream),
                    "sasvqa_git_vqa_generate_f32")
    out = out.long()
    if trim and B > 0:
        gen = out[:, L0:]
        is_eos = gen == eos_token_id
        first = torch.where(is_eos.any(dim=1), is_eos.float().argmax(dim=1) + 1, torch.full((B,), gen.shape[1], device=out.device))
        out = out[:, :L0 + int(first.max())]
    return out


def vqa_hidden(pixel_values: torch.Tensor, input_ids: torch.Tensor, enc: FrameEncoder, dec: GitDecoder, n_layers: int):
    """Inspection: (visual [B, K*197, 768], text [B, L, 768]) fp32 stream after ``n_layers`` decoder blocks."""
    px, ids = _prep(enc, pixel_values, input_ids)
    B, K = int(px.shape[0]), int(px.shape[1])
    L = int(ids.shape[1])
    nv = K * TOKENS
    out = torch.empty(B * nv + B * L, HIDDEN, dtype=torch.float32, device=enc.device)
    with torch.cuda.device(enc.device):
        _capi.check(_capi.lib().sasvqa_git_vqa_hidden_f32(dec.handle, enc.handle, px.data_ptr(), B, K, ids.data_ptr(), L,
                                                          int(n_layers), out.data_ptr(),
                                                          torch.cuda.current_stream().cuda_stream), "sasvqa_git_vqa_hidden_f32")
    return out[:B * nv].view(B, nv, HIDDEN), out[B * nv:].view(B, L, HIDDEN)
