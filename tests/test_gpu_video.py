"""Row f3 (pytest -m gpu): NVDEC decode front end in front of K0.  Known-answer test without any encoder in the image: a
lossless H.264 stream whose macroblocks are all I_PCM (tests/h264_pcm.py) must come back bit for bit; the kept-frame rule
is prefetch_loader.py:63 (`frame_count % intv == 0`); the RGB frames feed the sampler directly."""
import numpy as np
import pytest
import torch

import h264_pcm
import sasvqa_b200 as sas
from sasvqa_b200 import _capi, ops, synth

pytestmark = pytest.mark.gpu


def _frames(T, H, W, seed=0):
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    y = np.stack([np.clip(128 + 90 * np.sin(xx / 9.0 + t) * np.cos(yy / 7.0 - t / 3.0) + rng.randn(H, W) * 8, 0, 255) for t in range(T)])
    u = np.stack([np.clip(128 + 60 * np.sin(xx[::2, ::2] / 15.0 + t / 2.0) + rng.randn(H // 2, W // 2) * 4, 0, 255) for t in range(T)])
    v = np.stack([np.clip(128 + 60 * np.cos(yy[::2, ::2] / 11.0 - t / 2.0) + rng.randn(H // 2, W // 2) * 4, 0, 255) for t in range(T)])
    y[0, :16, :16] = 0           # runs of zero bytes: exercises the emulation-prevention bytes
    y[1] = 16                    # a flat frame
    return y.astype(np.uint8), u.astype(np.uint8), v.astype(np.uint8)


@pytest.fixture(scope="module")
def parser():
    """libnvcuvid's bitstream parser (host side of NVDEC): needs the driver's video library, not the engine."""
    torch.cuda.set_device(0)
    y, u, v = _frames(4, 48, 64)
    try:
        sas.probe_video(h264_pcm.encode_i_pcm(y, u, v))
    except _capi.SasvqaError as e:
        pytest.skip(f"libnvcuvid unavailable here: {e}")
    return True


@pytest.fixture(scope="module")
def nvdec(parser):
    """the decode engine itself: containers started without the `video` driver capability (NVIDIA_DRIVER_CAPABILITIES =
    compute,utility on this pool) get CUDA_ERROR_NO_DEVICE from cuvidCreateDecoder -- reported as a skip, never faked."""
    y, u, v = _frames(2, 48, 64)
    try:
        sas.decode_video(h264_pcm.encode_i_pcm(y, u, v))
    except _capi.SasvqaError as e:
        pytest.skip(f"NVDEC engine not exposed to this container: {e}")
    return True


@pytest.mark.parametrize("T,H,W,intv", [(7, 96, 128, 1), (10, 240, 320, 3), (5, 224, 224, 2), (33, 48, 64, 4)])
def test_probe_parses_the_stream_and_applies_the_intv_rule(parser, T, H, W, intv):
    """The real cuvid parser accepts the stream (sequence header -> frame size, one picture per access unit) and the
    kept-frame count follows prefetch_loader.py:63 (`frame_count % intv == 0`)."""
    y, u, v = _frames(T, H, W, seed=T)
    info = sas.probe_video(h264_pcm.encode_i_pcm(y, u, v), intv=intv)
    assert info == dict(width=W, height=H, frames=len(range(0, T, intv)), stream_frames=T)


@pytest.mark.parametrize("T,H,W,intv", [(7, 96, 128, 1), (10, 240, 320, 3), (5, 224, 224, 2)])
def test_lossless_stream_decodes_bit_exact(nvdec, T, H, W, intv):
    y, u, v = _frames(T, H, W, seed=H + W)
    stream = h264_pcm.encode_i_pcm(y, u, v)
    info = sas.probe_video(stream, intv=intv)
    kept = list(range(0, T, intv))                                  # prefetch_loader.py:63
    assert info == dict(width=W, height=H, frames=len(kept), stream_frames=T)
    nv12 = sas.decode_video(stream, intv=intv, nv12=True).cpu().numpy()
    assert nv12.shape == (len(kept), H * 3 // 2, W)
    np.testing.assert_array_equal(nv12[:, :H], y[kept])                                   # luma: bit exact
    np.testing.assert_array_equal(nv12[:, H:, 0::2].reshape(len(kept), H // 2, W // 2), u[kept])   # interleaved chroma
    np.testing.assert_array_equal(nv12[:, H:, 1::2].reshape(len(kept), H // 2, W // 2), v[kept])
    rgb = sas.decode_video(stream, intv=intv)
    assert rgb.is_cuda and rgb.dtype == torch.uint8 and tuple(rgb.shape) == (len(kept), H, W, 3)
    np.testing.assert_array_equal(rgb.cpu().numpy(), h264_pcm.yuv_to_rgb_bt601(y[kept], u[kept], v[kept]))


def test_decoded_frames_feed_the_sampler(nvdec):
    """decode -> K0 resize -> MDF on the device, against the same frames uploaded from the host."""
    T, H, W, K, Wn = 24, 240, 320, 4, 2
    y, u, v = _frames(T, H, W, seed=5)
    rgb = sas.decode_video(h264_pcm.encode_i_pcm(y, u, v))
    want = torch.from_numpy(h264_pcm.yuv_to_rgb_bt601(y, u, v)).cuda()
    enc = ops.FrameEncoder(synth.random_encoder_state_dict(synth.REF_SEED), chunk_frames=64)
    try:
        a = sas.sample_mdf_batch(rgb.unsqueeze(0), enc, K, Wn)
        b = sas.sample_mdf_batch(want.unsqueeze(0), enc, K, Wn)
        assert torch.equal(a["indices"], b["indices"]) and torch.equal(a["frames"], b["frames"])
    finally:
        enc.close()


def test_errors_are_reported_not_thrown(parser):
    with pytest.raises(_capi.SasvqaError):
        sas.probe_video(b"\x00\x00\x00\x01\x09\x10" * 4)            # access-unit delimiters only: no sequence header
    with pytest.raises(ValueError):
        sas.probe_video(b"")
