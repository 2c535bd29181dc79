"""Test infrastructure: a minimal H.264 (ITU-T H.264 / ISO 14496-10) Annex-B elementary-stream WRITER that codes every
macroblock as I_PCM -- raw 8-bit samples, no prediction, no transform, no entropy-coded residual -- so a conforming decoder
must return the input YUV 4:2:0 frames bit for bit.  That gives the NVDEC front end a known-answer test without any encoder
in the image (no ffmpeg / cv2 / x264 here).  Baseline profile, CAVLC, one IDR slice per frame, deblocking disabled.

Syntax written (clause numbers of the standard): NAL unit + emulation prevention 7.3.1, SPS 7.3.2.1.1, PPS 7.3.2.2, slice
header 7.3.3 (IDR, pic_order_cnt_type 2), macroblock_layer 7.3.5 with mb_type 25 = I_PCM and pcm samples.
"""
from __future__ import annotations

import numpy as np


class _Bits:
    def __init__(self):
        self.bytes = bytearray()
        self.cur, self.n = 0, 0

    def u(self, nbits: int, value: int):
        for i in range(nbits - 1, -1, -1):
            self.cur = (self.cur << 1) | ((value >> i) & 1)
            self.n += 1
            if self.n == 8:
                self.bytes.append(self.cur)
                self.cur, self.n = 0, 0

    def ue(self, v: int):                      # Exp-Golomb, 9.1
        v += 1
        nb = v.bit_length()
        self.u(nb - 1, 0)
        self.u(nb, v)

    def se(self, v: int):
        self.ue(2 * v - 1 if v > 0 else -2 * v)

    def align_zero(self):
        while self.n:
            self.u(1, 0)

    def raw(self, data: bytes):
        assert self.n == 0
        self.bytes += data

    def trailing(self):                        # rbsp_trailing_bits
        self.u(1, 1)
        self.align_zero()


def _nal(header: int, rbsp: bytes) -> bytes:
    """start code + NAL header + payload with emulation_prevention_three_byte inserted (7.4.1)."""
    arr = np.frombuffer(rbsp, dtype=np.uint8)
    out = bytearray(b"\x00\x00\x00\x01")
    out.append(header)
    zeros = 0
    # runs of zeros are rare in noise but common in flat frames: handle them exactly, byte by byte only near zero pairs
    cand = np.flatnonzero((arr[:-2] == 0) & (arr[1:-1] == 0) & (arr[2:] <= 3)) if arr.size >= 3 else np.array([], dtype=np.int64)
    if cand.size == 0 and not (arr.size >= 2 and arr[-1] == 0 and arr[-2] == 0):
        out += rbsp
        return bytes(out)
    for b in rbsp:
        if zeros >= 2 and b <= 3:
            out.append(3)
            zeros = 0
        out.append(b)
        zeros = zeros + 1 if b == 0 else 0
    return bytes(out)


def encode_i_pcm(y: np.ndarray, u: np.ndarray, v: np.ndarray) -> bytes:
    """y [T, H, W], u / v [T, H/2, W/2] uint8 (H, W multiples of 16) -> Annex-B H.264 stream, one IDR picture per frame."""
    T, H, W = y.shape
    assert H % 16 == 0 and W % 16 == 0 and u.shape == (T, H // 2, W // 2) == v.shape
    mbw, mbh = W // 16, H // 16
    sps = _Bits()
    sps.u(8, 66)                # profile_idc: Baseline
    sps.u(8, 0xC0)              # constraint_set0/1 flags
    sps.u(8, 40)                # level_idc 4.0
    sps.ue(0)                   # seq_parameter_set_id
    sps.ue(0)                   # log2_max_frame_num_minus4
    sps.ue(2)                   # pic_order_cnt_type 2: output order = decoding order
    sps.ue(1)                   # max_num_ref_frames
    sps.u(1, 0)                 # gaps_in_frame_num_value_allowed_flag
    sps.ue(mbw - 1)
    sps.ue(mbh - 1)
    sps.u(1, 1)                 # frame_mbs_only_flag
    sps.u(1, 1)                 # direct_8x8_inference_flag
    sps.u(1, 0)                 # frame_cropping_flag
    sps.u(1, 0)                 # vui_parameters_present_flag
    sps.trailing()
    pps = _Bits()
    pps.ue(0)                   # pic_parameter_set_id
    pps.ue(0)                   # seq_parameter_set_id
    pps.u(1, 0)                 # entropy_coding_mode_flag: CAVLC
    pps.u(1, 0)                 # bottom_field_pic_order_in_frame_present_flag
    pps.ue(0)                   # num_slice_groups_minus1
    pps.ue(0)                   # num_ref_idx_l0_default_active_minus1
    pps.ue(0)                   # num_ref_idx_l1_default_active_minus1
    pps.u(1, 0)                 # weighted_pred_flag
    pps.u(2, 0)                 # weighted_bipred_idc
    pps.se(0)                   # pic_init_qp_minus26
    pps.se(0)                   # pic_init_qs_minus26
    pps.se(0)                   # chroma_qp_index_offset
    pps.u(1, 1)                 # deblocking_filter_control_present_flag
    pps.u(1, 0)                 # constrained_intra_pred_flag
    pps.u(1, 0)                 # redundant_pic_cnt_present_flag
    pps.trailing()
    stream = bytearray(_nal(0x67, bytes(sps.bytes)) + _nal(0x68, bytes(pps.bytes)))
    for t in range(T):
        s = _Bits()
        s.ue(0)                 # first_mb_in_slice
        s.ue(7)                 # slice_type: I (all slices of the picture)
        s.ue(0)                 # pic_parameter_set_id
        s.u(4, 0)               # frame_num (IDR)
        s.ue(t & 0xFFFF)        # idr_pic_id: differs between consecutive IDR pictures
        s.u(1, 0)               # no_output_of_prior_pics_flag
        s.u(1, 0)               # long_term_reference_flag
        s.se(0)                 # slice_qp_delta
        s.ue(1)                 # disable_deblocking_filter_idc = 1
        for my in range(mbh):
            for mx in range(mbw):
                s.ue(25)        # mb_type I_PCM
                s.align_zero()  # pcm_alignment_zero_bit
                s.raw(y[t, my * 16:my * 16 + 16, mx * 16:mx * 16 + 16].tobytes())
                s.raw(u[t, my * 8:my * 8 + 8, mx * 8:mx * 8 + 8].tobytes())
                s.raw(v[t, my * 8:my * 8 + 8, mx * 8:mx * 8 + 8].tobytes())
        s.trailing()
        stream += _nal(0x65, bytes(s.bytes))
    return bytes(stream)


def yuv_to_rgb_bt601(y: np.ndarray, u: np.ndarray, v: np.ndarray) -> np.ndarray:
    """The library's colour step restated (BT.601 limited range, integer matrix, chroma of the co-sited 2x2 block):
    [T, H, W] + 2 x [T, H/2, W/2] uint8 -> [T, H, W, 3] uint8 RGB."""
    c = 298 * (y.astype(np.int32) - 16) + 128
    d = np.repeat(np.repeat(u.astype(np.int32) - 128, 2, axis=1), 2, axis=2)
    e = np.repeat(np.repeat(v.astype(np.int32) - 128, 2, axis=1), 2, axis=2)
    r = (c + 409 * e) >> 8
    g = (c - 100 * d - 208 * e) >> 8
    b = (c + 516 * d) >> 8
    return np.clip(np.stack([r, g, b], axis=-1), 0, 255).astype(np.uint8)
