"""sasvqa_b200 -- B200-native SAS-VQA frame-sampling hot path.

The package directory is ``sas-vqa_b200/``; import it as ``sasvqa_b200`` (alias shim at the
repo root).  The compute lives in ``libsasvqa_b200.so`` (hand-written sm_100a CUDA, C ABI in
``include/sasvqa.h``); this Python layer mirrors the reference's sampler interface and raises
if the library is missing -- there is no CPU fallback.
"""
from . import synth  # noqa: F401
from ._capi import LIB_PATH, SasvqaError  # noqa: F401
from .extract import generate_h5  # noqa: F401
from .ops import FrameEncoder, STATUS_EMPTY, STATUS_FALLBACK, STATUS_OK, STATUS_TOO_FEW  # noqa: F401
from .scorer import CaptionScorer, generate_inds  # noqa: F401
from .video import decode_video, probe_video  # noqa: F401
from .vqa import GitDecoder, vqa_generate, vqa_logits, vqa_loss  # noqa: F401
from .sampler import (  # noqa: F401
    encode_sampled_frames,
    mif_select,
    sample_frame_indices,
    sample_frames_uniform,
    sample_mdf_batch,
    sample_mdf_host,
    sample_mdf_ragged,
    sample_mif_batch,
    sample_mif_host,
    sample_representative_frames,
)

__all__ = [
    "CaptionScorer", "FrameEncoder", "decode_video", "probe_video", "GitDecoder", "vqa_generate", "vqa_logits", "vqa_loss", "SasvqaError", "generate_inds", "encode_sampled_frames", "generate_h5", "mif_select", "sample_frame_indices", "sample_frames_uniform",
    "sample_mdf_batch", "sample_mdf_host", "sample_mdf_ragged", "sample_mif_batch", "sample_mif_host", "sample_representative_frames", "synth",
]
