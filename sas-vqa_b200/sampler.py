"""Host-side mirror of the reference's sampler interface (same names, arguments, errors).

Reference call sites (src/preprocessing/extract_features.py:87-94):

    exted_frms = sample_representative_frames(video_frms, model, args.K, args.W, debug_counter)
    exted_frms = sample_frames_uniform(video_frms, K=args.K)
    exted_frms = sample_frame_indices(video_frms, args.K, 4, len(video_frms))

and the MIF expression ``scores[::ds_rate].topk(K)[1]`` (src/preprocessing/gen_sample.py:87-88).
All tensor work goes through libsasvqa_b200.so; there is no CPU implementation here.
"""
from __future__ import annotations

import weakref

import numpy as np
import torch

from . import ops
from ._capi import SasvqaError
from .synth import IMG

_ENCODERS: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def as_frame_encoder(model, chunk_frames: int = 0) -> ops.FrameEncoder:
    """``model`` may be a FrameEncoder, or the reference's own encoder object: an HF
    ``GitVisionModel`` (possibly inside ``DataParallel`` as at extract_features.py:48), whose
    frozen weights are uploaded once and cached per model object."""
    if isinstance(model, ops.FrameEncoder):
        return model
    inner = getattr(model, "module", model)
    if not hasattr(inner, "state_dict"):
        raise TypeError("model must be a sasvqa_b200.FrameEncoder or a GitVisionModel-compatible nn.Module")
    enc = _ENCODERS.get(inner)
    if enc is None:
        enc = ops.FrameEncoder(inner.state_dict(), chunk_frames=chunk_frames)   # raises if not ViT-B/16
        _ENCODERS[inner] = enc
    return enc


def sample_representative_frames(frames: torch.Tensor, model, K: int = 16, W: int = 8, debug_counter=None):
    """MDF sampler with the reference signature (src/preprocessing/datautils/utils.py:31-94).

    frames: (T, 3, 224, 224) fp32 normalised frames on any device.  Returns the K selected
    frames (K, 3, 224, 224), same dtype/device, in importance order (never index-sorted).
    ``debug_counter['Failure']`` / ``['Zeros']`` are bumped exactly where the reference does
    (and, like there, dereferenced unconditionally on those paths)."""
    enc = as_frame_encoder(model)
    T = int(frames.size(0))
    if T == 0:                                           # utils.py:50-52
        debug_counter["Zeros"] += 1
        return frames.new_zeros(K, 3, 224, 224)
    if frames.dim() != 4 or tuple(frames.shape[1:]) != (3, IMG, IMG):
        raise ValueError(f"Input image size ({frames.shape[-2]}*{frames.shape[-1]}) doesn't match model (224*224).")
    src = frames.to(device=enc.device, dtype=torch.float32).unsqueeze(0)
    res = ops.mdf_sample_device(enc, src, K, W, want_frames=True)
    status = int(res["status"][0])
    if status == ops.STATUS_TOO_FEW:                     # utils.py:92 -> torch.topk raises
        raise RuntimeError("selected index k out of range")
    if status == ops.STATUS_FALLBACK:                    # utils.py:93
        debug_counter["Failure"] += 1
    return res["frames"][0].to(device=frames.device, dtype=frames.dtype)


def sample_mdf_batch(clips: torch.Tensor, model, K: int = 16, W: int = 8, debug_counter=None, want_frames: bool = True,
                     want_aux: bool = False) -> dict:
    """Batched MDF over device-resident clips: [B, T, 224, 224, 3] uint8 (decoded RGB frames) or
    [B, T, 3, 224, 224] fp32.  Returns dict(indices int32 [B,K], status int32 [B], frames
    [B,K,3,224,224] fp32 | None, lcl_avg, feats).  Never raises for per-clip conditions: check
    ``status`` (ops.STATUS_*); counters are bumped if a dict is given."""
    enc = as_frame_encoder(model)
    res = ops.mdf_sample_device(enc, clips, K, W, want_frames=want_frames, want_aux=want_aux)
    if debug_counter is not None:
        st = res["status"].cpu()
        debug_counter["Failure"] += int((st == ops.STATUS_FALLBACK).sum())
        debug_counter["Zeros"] += int((st == ops.STATUS_EMPTY).sum())
    return res


def sample_mdf_ragged(clips, model, K: int = 16, W: int = 8, debug_counter=None, want_frames: bool = True,
                      want_aux: bool = False) -> dict:
    """Batched MDF over clips of DIFFERENT lengths (one frame size): ``clips`` is a sequence of uint8 tensors
    [T_i, H, W, 3] (GPU clips are concatenated on the device; CPU clips go through one pinned buffer and the
    pipelined host-buffer call, results then come back as host tensors).  One library call for the whole batch -- the
    encoder is frame-batched, the selection kernels take per-clip offsets -- with per clip exactly the result of the
    reference's one-video-per-call loop (extract_features.py:80-97): empty clips give zero frames ('Zeros'), W == -1
    adapts per clip, the fallback counts a 'Failure'.  Returns the dict of ``sample_mdf_batch`` with per-frame
    ``lcl_avg`` / ``feats`` packed, plus ``offsets``."""
    enc = as_frame_encoder(model)
    clips = [torch.as_tensor(c) for c in clips]
    lengths = [int(c.shape[0]) for c in clips]
    if not clips:
        raise ValueError("empty clip list")
    shapes = {tuple(c.shape[1:]) for c in clips}
    if len(shapes) != 1:
        raise ValueError(f"all clips of a ragged batch must share one frame size, got {sorted(shapes)}")
    if all(not c.is_cuda for c in clips) and not want_aux:
        # host clips: one pinned staging buffer, then the pipelined host-buffer call (results in host tensors)
        host = torch.empty((sum(lengths),) + tuple(clips[0].shape[1:]), dtype=torch.uint8, pin_memory=torch.cuda.is_available())
        torch.cat(clips, dim=0, out=host)
        res = ops.mdf_sample_ragged_host(enc, host, lengths, K, W, want_frames=want_frames)
    else:
        frames = torch.cat([c.to(enc.device) for c in clips], dim=0)
        res = ops.mdf_sample_ragged(enc, frames, lengths, K, W, want_frames=want_frames, want_aux=want_aux)
    if debug_counter is not None:
        st = res["status"].cpu()
        debug_counter["Failure"] += int((st == ops.STATUS_FALLBACK).sum())
        debug_counter["Zeros"] += int((st == ops.STATUS_EMPTY).sum())
    return res


def sample_mdf_host(clips_host: torch.Tensor, model, K: int = 16, W: int = 8, debug_counter=None, **kw) -> dict:
    """Same for clips in host memory ([B, T, 224, 224, 3] uint8, pinned for full speed): the
    extraction loop extract_features.py:80-97 over a clip list, copies overlapped with compute."""
    enc = as_frame_encoder(model)
    res = ops.mdf_sample_host(enc, clips_host, K, W, **kw)
    if debug_counter is not None:
        debug_counter["Failure"] += int((res["status"] == ops.STATUS_FALLBACK).sum())
        debug_counter["Zeros"] += int((res["status"] == ops.STATUS_EMPTY).sum())
    return res


def sample_frames_uniform(frames: torch.Tensor, K: int = 8) -> torch.Tensor:
    """utils.py:96-109: K frames at a truncating stride of T / K starting at int((T/K)//2)."""
    num_frames = len(frames)
    if num_frames <= K:
        print(num_frames)
    return frames[torch.as_tensor(uniform_indices(num_frames, K), dtype=torch.long, device=frames.device)]


def uniform_indices(T: int, K: int) -> list:
    intv = T / K
    cur = int(intv // 2)
    out = []
    for _ in range(K):
        out.append(cur)
        cur = int(cur + intv)
    return out


def sample_frame_indices(video_frms, clip_len: int, frame_sample_rate: int, seg_len: int):
    """extract_features.py:32-39 (GIT-6 baseline): random window, linspace, clip; numpy global RNG."""
    converted_len = int(clip_len * frame_sample_rate)
    end_idx = np.random.randint(converted_len, seg_len)
    start_idx = end_idx - converted_len
    assert start_idx >= 0
    indices = np.clip(np.linspace(start_idx, end_idx, num=clip_len), start_idx, end_idx - 1).astype(np.int64)
    return video_frms[indices]


def mif_select(scores: torch.Tensor, K: int, ds_rate: int = 1):
    """MIF index selection (gen_sample.py:87-88): top-K over every ds_rate-th relevance score,
    mapped back to frame indices, best first.  1-D scores -> python list (what the reference
    stores under 'sampled_inds'); 2-D [B, T] -> int32 tensor [B, K]."""
    s = scores.detach()
    if not s.is_cuda:
        if not torch.cuda.is_available():
            raise SasvqaError("mif_select needs a CUDA device (no CPU fallback)")
        s = s.cuda()
    idx = ops.topk_strided(s.float(), K, ds_rate)
    return idx.cpu().tolist() if scores.dim() == 1 else idx


def sample_mif_batch(clips: torch.Tensor, model, question_embeds: torch.Tensor, K: int = 8, ds_rate: int = 1,
                     want_frames: bool = False, want_aux: bool = False) -> dict:
    """Batched MIF with embedding-space relevance (BASELINE config 3): every frame of ``clips``
    ([B, T, H, W, 3] uint8 on the GPU) is encoded, scored against its clip's question embedding
    (``question_embeds`` [B, 768]) and the K best of every ``ds_rate``-th frame are returned best first --
    the selection rule of gen_sample.py:87-88.  dict(indices int32 [B, K], scores, feats, frames)."""
    enc = as_frame_encoder(model)
    return ops.mif_sample_device(enc, clips, question_embeds.to(device=clips.device, dtype=torch.float32), K, ds_rate,
                                 want_frames=want_frames, want_aux=want_aux)


def sample_mif_host(clips_host: torch.Tensor, model, question_embeds: torch.Tensor, K: int = 8, ds_rate: int = 1, **kw) -> dict:
    """``sample_mif_batch`` for clips and question embeddings in host memory (pinned for full speed): copies overlap
    compute in the library's double-buffered pipeline, the index table (and frames) come back in host tensors."""
    enc = as_frame_encoder(model)
    return ops.mif_sample_host(enc, clips_host, question_embeds, K, ds_rate, **kw)


def encode_sampled_frames(sampled: torch.Tensor, model, project: bool = True) -> torch.Tensor:
    """The visual side of the downstream video-QA forward (src/modeling/modeling.py:76-95,
    ``MyGitModel.forward`` with 5-D ``pixel_values``): ``sampled`` [B, K, 3, 224, 224] fp32 (what the collator
    builds from the ``sampled_frames`` rows) -> [B, K * 197, 768]: per-frame ``last_hidden_state`` concatenated
    along the sequence and passed through ``visual_projection`` (load it with ``FrameEncoder.set_projection``).
    All K frames of all clips run as one batch through the sampler's encoder kernels."""
    enc = as_frame_encoder(model)
    if sampled.dim() != 5:
        raise ValueError("pixel_values must be of rank 5: (batch_size, num_frames, num_channels, height, width)")
    tok = enc.visual_tokens(sampled.to(enc.device), project=project)          # [B, K, 197, 768]
    return tok.reshape(tok.shape[0], tok.shape[1] * tok.shape[2], tok.shape[3])
