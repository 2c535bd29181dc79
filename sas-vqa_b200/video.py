"""Decode front end: the reference's cv2 loop ``InputGen.__iter__`` (src/preprocessing/prefetch_loader.py:
50-76 -- ``cap.read()`` per frame, BGR -> RGB swap, keep frame i when ``i % intv == 0``) on the GPU's video engine.

``decode_video(bitstream, intv)`` takes an H.264 / HEVC Annex-B elementary stream (bytes in host memory; demux the dataset's
.avi / .mp4 on the host, e.g. ``ffmpeg -c copy -bsf:v h264_mp4toannexb``) and returns the kept frames as a uint8 CUDA tensor
``[T, H, W, 3]`` (RGB) -- exactly what ``sample_mdf_batch`` / ``sample_mdf_ragged`` / ``generate_h5`` take, so decoded frames
never visit host memory.  No CPU fallback: raises ``SasvqaError`` when libnvcuvid or an NVDEC engine is missing.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _capi

CODECS = {"h264": 4, "hevc": 8}


def _buffer(bitstream):
    arr = np.frombuffer(bitstream, dtype=np.uint8) if isinstance(bitstream, (bytes, bytearray, memoryview)) else \
        np.ascontiguousarray(bitstream, dtype=np.uint8)
    if arr.size == 0:
        raise ValueError("empty bitstream")
    return arr


def probe_video(bitstream, codec: str = "h264", intv: int = 1) -> dict:
    """dict(width, height, frames (what decode_video returns with this intv), stream_frames) -- parses, decodes nothing."""
    arr = _buffer(bitstream)
    info = np.zeros(4, dtype=np.int32)
    _capi.check(_capi.lib().sasvqa_video_probe(arr.ctypes.data, arr.size, CODECS[codec], int(intv), info.ctypes.data),
                "sasvqa_video_probe")
    return dict(width=int(info[0]), height=int(info[1]), frames=int(info[2]), stream_frames=int(info[3]))


def decode_video(bitstream, codec: str = "h264", intv: int = 1, device=None, nv12: bool = False) -> torch.Tensor:
    """All frames i of the stream with ``i % intv == 0``, display order, as uint8 ``[T, H, W, 3]`` RGB on the GPU
    (``nv12=True``: the decoder's raw output ``[T, H * 3 // 2, W]``)."""
    if not torch.cuda.is_available():
        raise _capi.SasvqaError("decode_video needs a CUDA device with an NVDEC engine (no CPU fallback)")
    arr = _buffer(bitstream)
    info = probe_video(arr, codec, intv)
    H, W, n = info["height"], info["width"], info["frames"]
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    shape = (n, H * 3 // 2, W) if nv12 else (n, H, W, 3)
    out = torch.empty(shape, dtype=torch.uint8, device=dev)
    got = np.zeros(1, dtype=np.int32)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().sasvqa_video_decode(arr.ctypes.data, arr.size, CODECS[codec], int(intv), out.data_ptr(), n, H, W,
                                                    int(bool(nv12)), got.ctypes.data, torch.cuda.current_stream(dev).cuda_stream),
                    "sasvqa_video_decode")
    return out[: int(got[0])]
