"""Torch-tensor wrappers over the C ABI, one per entry point of include/sasvqa.h.

Every function allocates outputs with torch (device memory plumbing only), passes raw pointers
and the current CUDA stream to libsasvqa_b200.so, and raises on any error.  Nothing here computes.
"""
from __future__ import annotations

import ctypes

import torch

from . import _capi
from .synth import HIDDEN, IMG, TOKENS, state_dict_keys

FRAME_ELEMS = 3 * IMG * IMG
PATCHES = (IMG // 16) ** 2
NUM_PARAMS = 85_799_424

STATUS_OK, STATUS_FALLBACK, STATUS_EMPTY, STATUS_TOO_FEW = 0, 1, 2, 3


def _need_cuda(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _capi.SasvqaError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    return t.contiguous()


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def flatten_state_dict(state_dict: dict) -> torch.Tensor:
    """fp32 CPU vector in the key order sasvqa_encoder_create expects (HF GitVisionModel order)."""
    parts = []
    for name, shape in state_dict_keys():
        if name not in state_dict:
            raise KeyError(f"encoder state dict lacks {name!r}; expected a GitVisionModel ViT-B/16 state dict")
        t = state_dict[name].detach().to("cpu", torch.float32)
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name}: shape {tuple(t.shape)} != {tuple(shape)} (only ViT-B/16 @224 is supported)")
        parts.append(t.reshape(-1))
    flat = torch.cat(parts).contiguous()
    assert flat.numel() == NUM_PARAMS
    return flat


class FrameEncoder:
    """Owns a SasvqaEncoder handle: bf16 weights, workspace and TMA descriptors on one GPU.
    The stand-in for `GitVisionModel.from_pretrained(...).eval().cuda()` (extract_features.py:145,45-47)."""

    def __init__(self, state_dict: dict, chunk_frames: int = 0, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        flat = flatten_state_dict(state_dict)
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            rc = _capi.lib().sasvqa_encoder_create(flat.data_ptr(), flat.numel(), int(chunk_frames),
                                                   ctypes.byref(handle))
        _capi.check(rc, "sasvqa_encoder_create")
        self._h = handle
        self.chunk_frames = _capi.lib().sasvqa_encoder_chunk_frames(self._h)

    @classmethod
    def from_model(cls, model, chunk_frames: int = 0, device=None) -> "FrameEncoder":
        """From an HF GitVisionModel (optionally wrapped in DataParallel like extract_features.py:48)."""
        inner = getattr(model, "module", model)
        return cls(inner.state_dict(), chunk_frames=chunk_frames, device=device)

    @property
    def handle(self):
        if self._h is None:
            raise _capi.SasvqaError("encoder handle already closed")
        return self._h

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            _capi.lib().sasvqa_encoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- instrumentation
    PROFILE_KINDS = ("preprocess", "gemm_patch_embed", "pre_layernorm", "layernorm", "gemm_qkv", "attention",
                     "gemm_out_proj", "gemm_fc1", "gemm_fc2", "pool_norm", "scores", "select", "gather", "resize",
                     "projection")

    def profile_enable(self, on: bool = True) -> None:
        _capi.check(_capi.lib().sasvqa_profile_enable(self.handle, int(on)), "sasvqa_profile_enable")

    def profile_read(self) -> dict:
        """{stage: (total_ms, n_scopes)} accumulated since the last read; synchronises the device."""
        n = len(self.PROFILE_KINDS)
        ms = (ctypes.c_double * n)()
        cnt = (ctypes.c_int64 * n)()
        _capi.check(_capi.lib().sasvqa_profile_read(self.handle, ms, cnt, n), "sasvqa_profile_read")
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.PROFILE_KINDS)}

    # ---- row f2: the downstream model's visual tokens on the sampled frames
    def set_projection(self, weight: torch.Tensor, bias: torch.Tensor, ln_weight: torch.Tensor, ln_bias: torch.Tensor):
        """Loads GIT's ``visual_projection`` (Linear(768, 768) + LayerNorm): ``visual_projection.0.weight``,
        ``.0.bias``, ``.1.weight``, ``.1.bias`` of the reference's ``MyGitModel`` (src/modeling/modeling.py:93)."""
        host = [t.detach().to("cpu", torch.float32).contiguous() for t in (weight, bias, ln_weight, ln_bias)]
        if tuple(host[0].shape) != (HIDDEN, HIDDEN) or any(tuple(t.shape) != (HIDDEN,) for t in host[1:]):
            raise ValueError("visual projection must be Linear(768, 768) + LayerNorm(768)")
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib().sasvqa_encoder_set_projection(self.handle, *[t.data_ptr() for t in host]),
                        "sasvqa_encoder_set_projection")

    def visual_tokens(self, frames: torch.Tensor, project: bool = True) -> torch.Tensor:
        """frames [..., 3, 224, 224] fp32 (processed frames, e.g. rows of ``sampled_frames``) or [..., 224, 224, 3]
        uint8 on the GPU -> [..., 197, 768] fp32: ``image_encoder(frame).last_hidden_state`` per frame, passed
        through ``visual_projection`` when ``project`` (src/modeling/modeling.py:76-95)."""
        if not frames.is_cuda:
            raise _capi.SasvqaError("frames must be a CUDA tensor (no CPU fallback)")
        frames = frames.contiguous()
        if frames.dtype == torch.float32 and tuple(frames.shape[-3:]) == (3, IMG, IMG):
            fn, name = _capi.lib().sasvqa_visual_tokens_f32, "sasvqa_visual_tokens_f32"
        elif frames.dtype == torch.uint8 and tuple(frames.shape[-3:]) == (IMG, IMG, 3):
            fn, name = _capi.lib().sasvqa_visual_tokens_u8, "sasvqa_visual_tokens_u8"
        else:
            raise ValueError(f"frames must be fp32 [..., 3, 224, 224] or uint8 [..., 224, 224, 3], got "
                             f"{frames.dtype} {tuple(frames.shape)}")
        lead = tuple(frames.shape[:-3])
        n = 1
        for d in lead:
            n *= d
        out = torch.empty(lead + (TOKENS, HIDDEN), dtype=torch.float32, device=frames.device)
        with torch.cuda.device(frames.device):
            _capi.check(fn(self.handle, frames.data_ptr(), n, int(bool(project)), out.data_ptr(), _stream(frames)), name)
        return out

    # ---- K2 + K3a
    def forward_patches(self, patches: torch.Tensor) -> torch.Tensor:
        patches = _need_cuda(patches, torch.bfloat16, "patches")
        n = patches.shape[0] // PATCHES
        feats = torch.empty(n, HIDDEN, dtype=torch.float32, device=patches.device)
        with torch.cuda.device(patches.device):
            _capi.check(_capi.lib().sasvqa_encoder_fwd(self.handle, patches.data_ptr(), n, feats.data_ptr(),
                                                       _stream(patches)), "sasvqa_encoder_fwd")
        return feats

    def hidden(self, patches: torch.Tensor, n_layers: int) -> torch.Tensor:
        patches = _need_cuda(patches, torch.bfloat16, "patches")
        n = patches.shape[0] // PATCHES
        out = torch.empty(n, TOKENS, HIDDEN, dtype=torch.float32, device=patches.device)
        with torch.cuda.device(patches.device):
            _capi.check(_capi.lib().sasvqa_encoder_fwd_hidden(self.handle, patches.data_ptr(), n, int(n_layers),
                                                              out.data_ptr(), _stream(patches)),
                        "sasvqa_encoder_fwd_hidden")
        return out

    def features_u8(self, frames_hwc: torch.Tensor) -> torch.Tensor:
        return self.forward_patches(preprocess_u8(frames_hwc))

    def features_f32(self, frames_chw: torch.Tensor) -> torch.Tensor:
        return self.forward_patches(patchify_f32(frames_chw))


# ---- K0
def resize_crop_u8(frames_hwc: torch.Tensor) -> torch.Tensor:
    """uint8 ``[..., H, W, 3]`` -> uint8 ``[..., 224, 224, 3]``: the image processor's shortest-edge bicubic
    resize + centre crop (``prefetch_loader.py:74-75``), bit-exact with the host implementation."""
    frames_hwc = _need_cuda(frames_hwc, torch.uint8, "frames")
    if frames_hwc.dim() < 3 or frames_hwc.shape[-1] != 3:
        raise ValueError(f"frames must be [..., H, W, 3], got {tuple(frames_hwc.shape)}")
    lead, (H, W) = tuple(frames_hwc.shape[:-3]), frames_hwc.shape[-3:-1]
    n = 1
    for d in lead:
        n *= d
    out = torch.empty(lead + (IMG, IMG, 3), dtype=torch.uint8, device=frames_hwc.device)
    with torch.cuda.device(frames_hwc.device):
        _capi.check(_capi.lib().sasvqa_resize_crop_u8(frames_hwc.data_ptr(), n, int(H), int(W), out.data_ptr(),
                                                      _stream(frames_hwc)), "sasvqa_resize_crop_u8")
    return out


# ---- K1
def preprocess_u8(frames_hwc: torch.Tensor) -> torch.Tensor:
    frames_hwc = _need_cuda(frames_hwc, torch.uint8, "frames")
    assert frames_hwc.shape[-3:] == (IMG, IMG, 3), "frames must be [..., 224, 224, 3] uint8"
    n = frames_hwc.numel() // FRAME_ELEMS
    patches = torch.empty(n * PATCHES, HIDDEN, dtype=torch.bfloat16, device=frames_hwc.device)
    with torch.cuda.device(frames_hwc.device):
        _capi.check(_capi.lib().sasvqa_preprocess_u8(frames_hwc.data_ptr(), n, patches.data_ptr(),
                                                     _stream(frames_hwc)), "sasvqa_preprocess_u8")
    return patches


def patchify_f32(frames_chw: torch.Tensor) -> torch.Tensor:
    frames_chw = _need_cuda(frames_chw, torch.float32, "frames")
    assert frames_chw.shape[-3:] == (3, IMG, IMG), "frames must be [..., 3, 224, 224] fp32"
    n = frames_chw.numel() // FRAME_ELEMS
    patches = torch.empty(n * PATCHES, HIDDEN, dtype=torch.bfloat16, device=frames_chw.device)
    with torch.cuda.device(frames_chw.device):
        _capi.check(_capi.lib().sasvqa_patchify_f32(frames_chw.data_ptr(), n, patches.data_ptr(),
                                                    _stream(frames_chw)), "sasvqa_patchify_f32")
    return patches


# ---- K3b + K4a
def mdf_scores(feats: torch.Tensor, W: int, want_gram: bool = False):
    feats = _need_cuda(feats, torch.float32, "feats")
    squeeze = feats.dim() == 2
    if squeeze:
        feats = feats.unsqueeze(0)
    B, T, D = feats.shape
    assert D == HIDDEN
    lcl = torch.empty(B, T, dtype=torch.float32, device=feats.device)
    gram = torch.empty(B, T, T, dtype=torch.float32, device=feats.device) if want_gram else None
    with torch.cuda.device(feats.device):
        _capi.check(_capi.lib().sasvqa_mdf_scores(feats.data_ptr(), B, T, int(W), lcl.data_ptr(), _capi.ptr(gram),
                                                  _stream(feats)), "sasvqa_mdf_scores")
    if squeeze:
        lcl = lcl[0]
        gram = gram[0] if gram is not None else None
    return (lcl, gram) if want_gram else lcl


# ---- K4b
def mdf_select(lcl_avg: torch.Tensor, K: int, W: int):
    lcl_avg = _need_cuda(lcl_avg, torch.float32, "lcl_avg")
    squeeze = lcl_avg.dim() == 1
    if squeeze:
        lcl_avg = lcl_avg.unsqueeze(0)
    B, T = lcl_avg.shape
    idx = torch.empty(B, K, dtype=torch.int32, device=lcl_avg.device)
    status = torch.empty(B, dtype=torch.int32, device=lcl_avg.device)
    with torch.cuda.device(lcl_avg.device):
        _capi.check(_capi.lib().sasvqa_mdf_select(lcl_avg.data_ptr(), B, T, int(K), int(W), idx.data_ptr(),
                                                  status.data_ptr(), _stream(lcl_avg)), "sasvqa_mdf_select")
    return (idx[0], status[0]) if squeeze else (idx, status)


# ---- K4c
def topk_strided(scores: torch.Tensor, K: int, ds_rate: int = 1) -> torch.Tensor:
    scores = _need_cuda(scores, torch.float32, "scores")
    squeeze = scores.dim() == 1
    if squeeze:
        scores = scores.unsqueeze(0)
    B, T = scores.shape
    idx = torch.empty(B, K, dtype=torch.int32, device=scores.device)
    with torch.cuda.device(scores.device):
        _capi.check(_capi.lib().sasvqa_topk_strided(scores.data_ptr(), B, T, int(ds_rate), int(K), idx.data_ptr(),
                                                    _stream(scores)), "sasvqa_topk_strided")
    return idx[0] if squeeze else idx


# ---- K5
def gather_frames_u8(clips: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    clips = _need_cuda(clips, torch.uint8, "clips")
    idx = _need_cuda(idx, torch.int32, "idx")
    B, T = clips.shape[0], clips.shape[1]
    K = idx.shape[-1]
    out = torch.empty(B, K, 3, IMG, IMG, dtype=torch.float32, device=clips.device)
    with torch.cuda.device(clips.device):
        _capi.check(_capi.lib().sasvqa_gather_frames_u8(clips.data_ptr(), idx.data_ptr(), B, T, K, out.data_ptr(),
                                                        _stream(clips)), "sasvqa_gather_frames_u8")
    return out


def gather_frames_f32(frames: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """frames [B, T, ...] fp32, idx [B, K] -> [B, K, ...]."""
    frames = _need_cuda(frames, torch.float32, "frames")
    idx = _need_cuda(idx, torch.int32, "idx")
    B, T = frames.shape[0], frames.shape[1]
    K = idx.shape[-1]
    row = frames[0, 0].numel() if T > 0 else 0
    out = torch.empty((B, K) + tuple(frames.shape[2:]), dtype=torch.float32, device=frames.device)
    with torch.cuda.device(frames.device):
        _capi.check(_capi.lib().sasvqa_gather_frames_f32(frames.data_ptr(), idx.data_ptr(), B, T, K, row,
                                                         out.data_ptr(), _stream(frames)), "sasvqa_gather_frames_f32")
    return out


# ---- whole path
def mdf_sample_device(enc: FrameEncoder, clips: torch.Tensor, K: int, W: int, want_frames: bool = True,
                      want_aux: bool = False) -> dict:
    """clips: [B, T, H, W, 3] uint8 (any frame size) or [B, T, 3, 224, 224] fp32, on the GPU."""
    if not clips.is_cuda:
        raise _capi.SasvqaError("clips must be a CUDA tensor (use mdf_sample_host for host buffers)")
    clips = clips.contiguous()
    B, T = clips.shape[0], clips.shape[1]
    dev = clips.device
    idx = torch.empty(B, K, dtype=torch.int32, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    lcl = torch.empty(B, T, dtype=torch.float32, device=dev) if want_aux else None
    feats = torch.empty(B, T, HIDDEN, dtype=torch.float32, device=dev) if want_aux else None
    sampled = torch.empty(B, K, 3, IMG, IMG, dtype=torch.float32, device=dev) if want_frames else None
    if clips.dtype == torch.uint8:
        if clips.dim() != 5 or clips.shape[-1] != 3:
            raise ValueError(f"uint8 clips must be [B, T, H, W, 3], got {tuple(clips.shape)}")
        H, Wd = int(clips.shape[2]), int(clips.shape[3])
        with torch.cuda.device(dev):
            _capi.check(_capi.lib().sasvqa_mdf_sample_u8_hw(
                enc.handle, clips.data_ptr(), B, T, H, Wd, int(K), int(W), idx.data_ptr(), status.data_ptr(),
                _capi.ptr(lcl), _capi.ptr(feats), _capi.ptr(sampled), _stream(clips)), "sasvqa_mdf_sample_u8_hw")
    elif clips.dtype == torch.float32:
        if tuple(clips.shape[2:]) != (3, IMG, IMG):
            raise ValueError(f"fp32 clips must be processed frames [B, T, 3, 224, 224], got {tuple(clips.shape)}")
        with torch.cuda.device(dev):
            _capi.check(_capi.lib().sasvqa_mdf_sample_f32(
                enc.handle, clips.data_ptr(), B, T, int(K), int(W), idx.data_ptr(), status.data_ptr(),
                _capi.ptr(lcl), _capi.ptr(feats), _capi.ptr(sampled), _stream(clips)), "sasvqa_mdf_sample_f32")
    else:
        raise TypeError(f"clips must be uint8 HWC or float32 CHW, got {clips.dtype}")
    return dict(indices=idx, status=status, lcl_avg=lcl, feats=feats, frames=sampled)


def mdf_sample_ragged(enc: FrameEncoder, frames: torch.Tensor, lengths, K: int, W: int, want_frames: bool = True,
                      want_aux: bool = False) -> dict:
    """Ragged batch: ``frames`` [sum(lengths), H, W, 3] uint8 on the GPU, clip b = the next ``lengths[b]`` frames.
    Returns dict(indices int32 [B, K], status [B], frames [B, K, 3, 224, 224] | None, lcl_avg [sum T], feats [sum T, 768],
    offsets int32 [B + 1] on the CPU)."""
    frames = _need_cuda(frames, torch.uint8, "frames")
    if frames.dim() != 4 or frames.shape[-1] != 3:
        raise ValueError(f"frames must be [sum T, H, W, 3], got {tuple(frames.shape)}")
    lens = torch.as_tensor(lengths, dtype=torch.int64).reshape(-1)
    if int(lens.sum()) != int(frames.shape[0]) or bool((lens < 0).any()):
        raise ValueError(f"lengths sum to {int(lens.sum())}, frames holds {int(frames.shape[0])}")
    B = int(lens.numel())
    off = torch.zeros(B + 1, dtype=torch.int32)
    off[1:] = torch.cumsum(lens, 0).to(torch.int32)
    nf, H, Wd = (int(v) for v in frames.shape[:3])
    dev = frames.device
    idx = torch.empty(B, K, dtype=torch.int32, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    lcl = torch.empty(nf, dtype=torch.float32, device=dev) if want_aux else None
    feats = torch.empty(nf, HIDDEN, dtype=torch.float32, device=dev) if want_aux else None
    sampled = torch.empty(B, K, 3, IMG, IMG, dtype=torch.float32, device=dev) if want_frames else None
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().sasvqa_mdf_sample_ragged_u8(
            enc.handle, frames.data_ptr(), B, off.data_ptr(), H, Wd, int(K), int(W), idx.data_ptr(), status.data_ptr(),
            _capi.ptr(lcl), _capi.ptr(feats), _capi.ptr(sampled), _stream(frames)), "sasvqa_mdf_sample_ragged_u8")
    return dict(indices=idx, status=status, lcl_avg=lcl, feats=feats, frames=sampled, offsets=off)


def mdf_sample_ragged_host(enc: FrameEncoder, frames_host: torch.Tensor, lengths, K: int, W: int, want_frames: bool = True,
                           idx_out: torch.Tensor = None, status_out: torch.Tensor = None, frames_out: torch.Tensor = None) -> dict:
    """Ragged batch from host memory: ``frames_host`` [sum(lengths), H, W, 3] uint8 (pinned for full speed).  Results
    land in (pinned) host tensors: dict(indices [B, K], status [B], frames [B, K, 3, 224, 224] | None, offsets); pass
    ``*_out`` to reuse pinned result buffers across calls (pinning 2.4 GB costs more than sampling 256 clips)."""
    if frames_host.is_cuda or frames_host.dtype != torch.uint8:
        raise TypeError("frames_host must be a uint8 CPU tensor")
    if frames_host.dim() != 4 or frames_host.shape[-1] != 3:
        raise ValueError(f"frames_host must be [sum T, H, W, 3], got {tuple(frames_host.shape)}")
    frames_host = frames_host.contiguous()
    lens = torch.as_tensor(lengths, dtype=torch.int64).reshape(-1)
    if int(lens.sum()) != int(frames_host.shape[0]) or bool((lens < 0).any()):
        raise ValueError(f"lengths sum to {int(lens.sum())}, frames_host holds {int(frames_host.shape[0])}")
    B = int(lens.numel())
    off = torch.zeros(B + 1, dtype=torch.int32)
    off[1:] = torch.cumsum(lens, 0).to(torch.int32)
    H, Wd = int(frames_host.shape[1]), int(frames_host.shape[2])
    pin = torch.cuda.is_available()
    idx = idx_out if idx_out is not None else torch.empty(B, K, dtype=torch.int32, pin_memory=pin)
    status = status_out if status_out is not None else torch.empty(B, dtype=torch.int32, pin_memory=pin)
    sampled = None
    if want_frames:
        sampled = frames_out if frames_out is not None else torch.empty(B, K, 3, IMG, IMG, dtype=torch.float32, pin_memory=pin)
    for t, shape, dt in ((idx, (B, K), torch.int32), (status, (B,), torch.int32), (sampled, (B, K, 3, IMG, IMG), torch.float32)):
        if t is not None and (tuple(t.shape) != shape or t.dtype != dt or t.is_cuda or not t.is_contiguous()):
            raise ValueError(f"result buffer must be a contiguous CPU {dt} tensor of shape {shape}")
    with torch.cuda.device(enc.device):
        _capi.check(_capi.lib().sasvqa_mdf_sample_ragged_host(enc.handle, frames_host.data_ptr(), B, off.data_ptr(), H, Wd, int(K),
                                                              int(W), idx.data_ptr(), status.data_ptr(), _capi.ptr(sampled)),
                    "sasvqa_mdf_sample_ragged_host")
    return dict(indices=idx, status=status, frames=sampled, offsets=off)


def mdf_sample_host(enc: FrameEncoder, clips_host: torch.Tensor, K: int, W: int, idx_out: torch.Tensor = None,
                    status_out: torch.Tensor = None, frames_out: torch.Tensor = None, want_frames: bool = True) -> dict:
    """clips_host: [B, T, H, W, 3] uint8 in (ideally pinned) host memory.  Results land in host tensors."""
    if clips_host.is_cuda or clips_host.dtype != torch.uint8:
        raise TypeError("clips_host must be a uint8 CPU tensor")
    if clips_host.dim() != 5 or clips_host.shape[-1] != 3:
        raise ValueError(f"clips_host must be [B, T, H, W, 3], got {tuple(clips_host.shape)}")
    clips_host = clips_host.contiguous()
    B, T, H, Wd = (int(v) for v in clips_host.shape[:4])
    pin = torch.cuda.is_available()
    if idx_out is None:
        idx_out = torch.empty(B, K, dtype=torch.int32, pin_memory=pin)
    if status_out is None:
        status_out = torch.empty(B, dtype=torch.int32, pin_memory=pin)
    if frames_out is None and want_frames:
        frames_out = torch.empty(B, K, 3, IMG, IMG, dtype=torch.float32, pin_memory=pin)
    with torch.cuda.device(enc.device):
        _capi.check(_capi.lib().sasvqa_mdf_sample_host_hw(enc.handle, clips_host.data_ptr(), B, T, H, Wd, int(K), int(W),
                                                          idx_out.data_ptr(), status_out.data_ptr(),
                                                          _capi.ptr(frames_out) if want_frames else None),
                    "sasvqa_mdf_sample_host_hw")
    return dict(indices=idx_out, status=status_out, frames=frames_out if want_frames else None)


def mif_scores(feats: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """feats [B, T, 768] fp32, q [B, 768] fp32 -> relevance scores [B, T] = <feats[b, t], q[b]>."""
    feats = _need_cuda(feats, torch.float32, "feats")
    q = _need_cuda(q, torch.float32, "q")
    B, T = feats.shape[0], feats.shape[1]
    if tuple(q.shape) != (B, HIDDEN) or feats.shape[2] != HIDDEN:
        raise ValueError(f"need feats [B, T, 768] and q [B, 768], got {tuple(feats.shape)} and {tuple(q.shape)}")
    out = torch.empty(B, T, dtype=torch.float32, device=feats.device)
    with torch.cuda.device(feats.device):
        _capi.check(_capi.lib().sasvqa_mif_scores(feats.data_ptr(), q.data_ptr(), B, T, out.data_ptr(), _stream(feats)),
                    "sasvqa_mif_scores")
    return out


def mif_sample_device(enc: FrameEncoder, clips: torch.Tensor, q: torch.Tensor, K: int, ds_rate: int = 1,
                      want_frames: bool = False, want_aux: bool = False) -> dict:
    """clips [B, T, H, W, 3] uint8 on the GPU, q [B, 768] fp32 question embeddings."""
    clips = _need_cuda(clips, torch.uint8, "clips")
    q = _need_cuda(q, torch.float32, "q")
    if clips.dim() != 5 or clips.shape[-1] != 3:
        raise ValueError(f"clips must be [B, T, H, W, 3], got {tuple(clips.shape)}")
    B, T, H, Wd = (int(v) for v in clips.shape[:4])
    if tuple(q.shape) != (B, HIDDEN):
        raise ValueError(f"q must be [B, 768], got {tuple(q.shape)}")
    dev = clips.device
    idx = torch.empty(B, K, dtype=torch.int32, device=dev)
    scores = torch.empty(B, T, dtype=torch.float32, device=dev) if want_aux else None
    feats = torch.empty(B, T, HIDDEN, dtype=torch.float32, device=dev) if want_aux else None
    sampled = torch.empty(B, K, 3, IMG, IMG, dtype=torch.float32, device=dev) if want_frames else None
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().sasvqa_mif_sample_u8_hw(
            enc.handle, clips.data_ptr(), B, T, H, Wd, q.data_ptr(), int(K), int(ds_rate), idx.data_ptr(),
            _capi.ptr(scores), _capi.ptr(feats), _capi.ptr(sampled), _stream(clips)), "sasvqa_mif_sample_u8_hw")
    return dict(indices=idx, scores=scores, feats=feats, frames=sampled)


def mif_sample_host(enc: FrameEncoder, clips_host: torch.Tensor, q_host: torch.Tensor, K: int, ds_rate: int = 1,
                    idx_out: torch.Tensor = None, frames_out: torch.Tensor = None, want_frames: bool = False) -> dict:
    """MIF from host memory: clips_host [B, T, H, W, 3] uint8 and q_host [B, 768] fp32 (pinned for full speed); the
    index table (and the sampled frames) land in host tensors."""
    if clips_host.is_cuda or clips_host.dtype != torch.uint8 or clips_host.dim() != 5 or clips_host.shape[-1] != 3:
        raise TypeError("clips_host must be a uint8 CPU tensor [B, T, H, W, 3]")
    if q_host.is_cuda or q_host.dtype != torch.float32:
        raise TypeError("q_host must be a float32 CPU tensor")
    clips_host, q_host = clips_host.contiguous(), q_host.contiguous()
    B, T, H, Wd = (int(v) for v in clips_host.shape[:4])
    if tuple(q_host.shape) != (B, HIDDEN):
        raise ValueError(f"q_host must be [B, 768], got {tuple(q_host.shape)}")
    pin = torch.cuda.is_available()
    if idx_out is None:
        idx_out = torch.empty(B, K, dtype=torch.int32, pin_memory=pin)
    if frames_out is None and want_frames:
        frames_out = torch.empty(B, K, 3, IMG, IMG, dtype=torch.float32, pin_memory=pin)
    with torch.cuda.device(enc.device):
        _capi.check(_capi.lib().sasvqa_mif_sample_host_hw(enc.handle, clips_host.data_ptr(), B, T, H, Wd, q_host.data_ptr(),
                                                          int(K), int(ds_rate), idx_out.data_ptr(),
                                                          _capi.ptr(frames_out) if want_frames else None),
                    "sasvqa_mif_sample_host_hw")
    return dict(indices=idx_out, frames=frames_out if want_frames else None)


def launch_count() -> int:
    return int(_capi.lib().sasvqa_launch_count())


# ---- test hooks
def test_gemm(a: torch.Tensor, b: torch.Tensor, mode: int, vec: torch.Tensor, out_f32: torch.Tensor = None,
              use_simt: bool = False):
    """One GEMM with a fused epilogue.  ``use_simt`` runs the CUDA-core check kernel of the TEST-ONLY library
    (libsasvqa_b200_test.so) instead of the product's tcgen05 kernel."""
    a = _need_cuda(a, torch.bfloat16, "a")
    b = _need_cuda(b, torch.bfloat16, "b")
    M, K = a.shape
    N = b.shape[0]
    out_bf16 = torch.empty(M, N, dtype=torch.bfloat16, device=a.device) if mode in (0, 1, 4) else None
    with torch.cuda.device(a.device):
        if use_simt:
            _capi.check_test(_capi.test_lib().sasvqa_check_gemm_simt(
                a.data_ptr(), b.data_ptr(), M, N, K, mode, vec.data_ptr(), _capi.ptr(out_bf16), _capi.ptr(out_f32),
                _stream(a)), "sasvqa_check_gemm_simt")
        else:
            _capi.check(_capi.lib().sasvqa_test_gemm(a.data_ptr(), b.data_ptr(), M, N, K, mode, vec.data_ptr(),
                                                     _capi.ptr(out_bf16), _capi.ptr(out_f32), _stream(a)),
                        "sasvqa_test_gemm")
    return out_bf16 if mode in (0, 1, 4) else out_f32


def test_attention(qkv: torch.Tensor, impl: int = 0) -> torch.Tensor:
    """The encoder's per-frame attention: impl 0 = the product's tcgen05 kernel, 1 = the mma.sync check kernel of the
    test-only library, 10 + v = instrumented variant v of the tcgen05 kernel."""
    qkv = _need_cuda(qkv, torch.bfloat16, "qkv")
    n = qkv.shape[0] // TOKENS
    out = torch.empty(n * TOKENS, HIDDEN, dtype=torch.bfloat16, device=qkv.device)
    with torch.cuda.device(qkv.device):
        if impl == 1:
            _capi.check_test(_capi.test_lib().sasvqa_check_attention_mma(qkv.data_ptr(), n, out.data_ptr(), _stream(qkv)),
                             "sasvqa_check_attention_mma")
        else:
            _capi.check(_capi.lib().sasvqa_test_attention(qkv.data_ptr(), n, out.data_ptr(), max(0, int(impl) - 10),
                                                          _stream(qkv)), "sasvqa_test_attention")
    return out


def test_layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor) -> torch.Tensor:
    x = _need_cuda(x, torch.float32, "x")
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _capi.check(_capi.lib().sasvqa_test_layernorm(x.data_ptr(), x.shape[0], gamma.data_ptr(), beta.data_ptr(),
                                                      out.data_ptr(), _stream(x)), "sasvqa_test_layernorm")
    return out
