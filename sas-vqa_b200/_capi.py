"""ctypes binding of libsasvqa_b200.so (C ABI in include/sasvqa.h).

There is no CPU fallback: if the library is missing or a call fails this raises.  PyTorch is
used only for device memory and streams; every signature below is plain pointers and sizes.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_uint8, c_uint16, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsasvqa_b200%s.so" % os.environ.get("SASVQA_LIB_SUFFIX", ""))

_p = c_void_p  # device / host pointers are passed as raw addresses

SIGNATURES = {
    "sasvqa_abi_version": (c_int, []),
    "sasvqa_last_error": (c_char_p, []),
    "sasvqa_encoder_create": (c_int, [_p, c_uint64, c_int, POINTER(c_void_p)]),
    "sasvqa_encoder_destroy": (None, [_p]),
    "sasvqa_encoder_chunk_frames": (c_int, [_p]),
    "sasvqa_resize_crop_u8": (c_int, [_p, c_int, c_int, c_int, _p, _p]),
    "sasvqa_preprocess_u8": (c_int, [_p, c_int, _p, _p]),
    "sasvqa_patchify_f32": (c_int, [_p, c_int, _p, _p]),
    "sasvqa_encoder_fwd": (c_int, [_p, _p, c_int, _p, _p]),
    "sasvqa_encoder_fwd_hidden": (c_int, [_p, _p, c_int, c_int, _p, _p]),
    "sasvqa_mdf_scores": (c_int, [_p, c_int, c_int, c_int, _p, _p, _p]),
    "sasvqa_mdf_select": (c_int, [_p, c_int, c_int, c_int, c_int, _p, _p, _p]),
    "sasvqa_topk_strided": (c_int, [_p, c_int, c_int, c_int, c_int, _p, _p]),
    "sasvqa_encoder_set_projection": (c_int, [_p, _p, _p, _p, _p]),
    "sasvqa_visual_tokens_f32": (c_int, [_p, _p, c_int, c_int, _p, _p]),
    "sasvqa_visual_tokens_u8": (c_int, [_p, _p, c_int, c_int, _p, _p]),
    "sasvqa_mif_scores": (c_int, [_p, _p, c_int, c_int, _p, _p]),
    "sasvqa_mif_sample_u8_hw": (c_int, [_p, _p, c_int, c_int, c_int, c_int, _p, c_int, c_int, _p, _p, _p, _p, _p]),
    "sasvqa_mif_sample_host_hw": (c_int, [_p, _p, c_int, c_int, c_int, c_int, _p, c_int, c_int, _p, _p]),
    "sasvqa_gather_frames_u8": (c_int, [_p, _p, c_int, c_int, c_int, _p, _p]),
    "sasvqa_gather_frames_f32": (c_int, [_p, _p, c_int, c_int, c_int, c_int64, _p, _p]),
    "sasvqa_mdf_sample_u8": (c_int, [_p, _p, c_int, c_int, c_int, c_int, _p, _p, _p, _p, _p, _p]),
    "sasvqa_mdf_sample_f32": (c_int, [_p, _p, c_int, c_int, c_int, c_int, _p, _p, _p, _p, _p, _p]),
    "sasvqa_mdf_sample_u8_hw": (c_int, [_p, _p, c_int, c_int, c_int, c_int, c_int, c_int, _p, _p, _p, _p, _p, _p]),
    "sasvqa_mdf_sample_ragged_u8": (c_int, [_p, _p, c_int, _p, c_int, c_int, c_int, c_int, _p, _p, _p, _p, _p, _p]),
    "sasvqa_mdf_sample_ragged_host": (c_int, [_p, _p, c_int, _p, c_int, c_int, c_int, c_int, _p, _p, _p]),
    "sasvqa_mdf_sample_host": (c_int, [_p, _p, c_int, c_int, c_int, c_int, _p, _p, _p]),
    "sasvqa_mdf_sample_host_hw": (c_int, [_p, _p, c_int, c_int, c_int, c_int, c_int, c_int, _p, _p, _p]),
    "sasvqa_scorer_num_params": (c_uint64, [c_int, c_int]),
    "sasvqa_scorer_create": (c_int, [_p, c_uint64, c_int, c_int, c_int, POINTER(c_void_p)]),
    "sasvqa_scorer_destroy": (None, [_p]),
    "sasvqa_scorer_max_tokens": (c_int, [_p]),
    "sasvqa_scorer_logits": (c_int, [_p, _p, _p, _p, c_int, c_int, _p, _p]),
    "sasvqa_scorer_hidden": (c_int, [_p, _p, _p, _p, c_int, c_int, c_int, _p, _p]),
    "sasvqa_scorer_logits_host": (c_int, [_p, _p, _p, _p, c_int, c_int, _p]),
    "sasvqa_mif_select_captions_host": (c_int, [_p, _p, _p, _p, c_int, c_int, c_int, c_int, c_int, c_int, _p, _p]),
    "sasvqa_scorer_profile_enable": (c_int, [_p, c_int]),
    "sasvqa_scorer_profile_read": (c_int, [_p, POINTER(ctypes.c_double), POINTER(c_int64), c_int]),
    "sasvqa_git_decoder_num_params": (c_uint64, [c_int, c_int]),
    "sasvqa_git_decoder_create": (c_int, [_p, c_uint64, c_int, c_int, c_int, POINTER(c_void_p)]),
    "sasvqa_git_decoder_destroy": (None, [_p]),
    "sasvqa_git_decoder_vocab_padded": (c_int, [_p]),
    "sasvqa_git_vqa_logits_f32": (c_int, [_p, _p, _p, c_int, c_int, _p, c_int, _p, _p]),
    "sasvqa_git_vqa_loss_f32": (c_int, [_p, _p, _p, c_int, c_int, _p, _p, c_int, _p, _p, _p]),
    "sasvqa_git_vqa_generate_f32": (c_int, [_p, _p, _p, c_int, c_int, _p, c_int, c_int, c_int, c_int, _p, _p]),
    "sasvqa_git_vqa_hidden_f32": (c_int, [_p, _p, _p, c_int, c_int, _p, c_int, c_int, _p, _p]),
    "sasvqa_test_attention_git": (c_int, [_p, c_int, c_int, c_int, _p, _p]),
    "sasvqa_test_attention_varlen": (c_int, [_p, _p, c_int, c_int, _p, _p]),
    "sasvqa_video_probe": (c_int, [_p, c_uint64, c_int, c_int, _p]),
    "sasvqa_video_decode": (c_int, [_p, c_uint64, c_int, c_int, _p, c_int, c_int, c_int, c_int, _p, _p]),
    "sasvqa_launch_count": (c_int64, []),
    "sasvqa_profile_enable": (c_int, [_p, c_int]),
    "sasvqa_profile_read": (c_int, [_p, POINTER(ctypes.c_double), POINTER(c_int64), c_int]),
    "sasvqa_test_gemm": (c_int, [_p, _p, c_int, c_int, c_int, c_int, _p, _p, _p, _p]),
    "sasvqa_test_attention": (c_int, [_p, c_int, _p, c_int, _p]),
    "sasvqa_test_layernorm": (c_int, [_p, c_int, _p, _p, _p, _p]),
}

# test-only library (csrc/check/): CUDA-core GEMM and mma.sync attention check kernels, never loaded by the product path
TEST_LIB_PATH = os.path.join(HERE, "libsasvqa_b200_test.so")
TEST_SIGNATURES = {
    "sasvqa_check_last_error": (c_char_p, []),
    "sasvqa_check_gemm_simt": (c_int, [_p, _p, c_int, c_int, c_int, c_int, _p, _p, _p, _p]),
    "sasvqa_check_attention_mma": (c_int, [_p, c_int, _p, _p]),
}

_lib = None
_test_lib = None


class SasvqaError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Loads the CUDA library; raises loudly when it has not been built (python -m sasvqa_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SasvqaError(
                f"{LIB_PATH} is missing: the CUDA extension is not built and there is no CPU fallback. "
                "Run `python __graft_entry__.py` (or sasvqa_b200.build.build()).")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def test_lib() -> ctypes.CDLL:
    """The check kernels the parity tests compare the product kernels against (tests and tools only)."""
    global _test_lib
    if _test_lib is None:
        if not os.path.exists(TEST_LIB_PATH):
            raise SasvqaError(f"{TEST_LIB_PATH} is missing: run `python __graft_entry__.py`")
        handle = ctypes.CDLL(TEST_LIB_PATH)
        for name, (res, args) in TEST_SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _test_lib = handle
    return _test_lib


def check_test(rc: int, what: str) -> None:
    if rc != 0:
        msg = test_lib().sasvqa_check_last_error()
        raise SasvqaError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().sasvqa_last_error()
        raise SasvqaError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def ptr(t) -> int | None:
    """Raw address of a torch tensor (device or host) or None."""
    if t is None:
        return None
    assert t.is_contiguous(), "tensor must be contiguous"
    return t.data_ptr()


def current_stream_ptr(device=None) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream
