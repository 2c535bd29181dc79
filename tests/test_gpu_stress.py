"""Parity under stress (pytest -m gpu): the cases benign random weights and mid-sized clips never reach.

* encoder with MASSIVE residual channels, heavy-tailed LayerNorm gains and weights that are not bf16-representable
  (what real CLIP / GIT checkpoints look like) against the fp32 oracle;
* clips whose scores tie exactly (all frames identical; two scenes of identical frames) -- utils.py:57-93 on ties;
* extreme lengths end to end: T in {1, 2, 3} and T in {4097, 10 000} with K = 2048 (the selection-rounds path, > 4096
  candidates), checked stage-wise against the restated sampler (oracle/mdf.py == utils.py:31-94);
* the handle serialises its own work across streams: an asynchronous device call on a side stream followed, with no
  synchronisation, by the host-buffer pipeline on the same encoder.
"""
import os

import pytest
import torch

from oracle import mdf, vit
import sasvqa_b200 as sas
from sasvqa_b200 import ops, synth

pytestmark = pytest.mark.gpu


class _FeatureModel:
    """Oracle-side stand-in for the encoder: returns given per-frame features as a 1-token last_hidden_state, chunk by
    chunk in the order utils.py:37-40 asks for them (mean over one token = the feature)."""

    def __init__(self, feats):
        self.feats, self.cursor = feats, 0

    def __call__(self, frames):
        n = frames.shape[0]
        out = self.feats[self.cursor:self.cursor + n]
        self.cursor = (self.cursor + n) % self.feats.shape[0]
        return type("O", (), {"last_hidden_state": out.unsqueeze(1)})()


def _excused(got, want, lcl_ref, eps):
    """picks equal, or the oracle's own scores of the two picks within 2 * eps (the north star's tie rule)"""
    bad = [(a, b) for a, b in zip(got, want) if a != b and abs(float(lcl_ref[a]) - float(lcl_ref[b])) > 2 * eps]
    return bad


@pytest.fixture(scope="module")
def encoder():
    torch.cuda.set_device(0)
    enc = ops.FrameEncoder(synth.random_encoder_state_dict(synth.REF_SEED), chunk_frames=512)
    yield enc
    enc.close()


def test_encoder_with_outlier_channels_and_unrounded_weights():
    """massive channels in the fp32 residual -> bf16 `h`.  Feature cosine >= 0.9999 per frame, windowed
    scores within 1e-3, and the picks identical up to tolerance-excused ties -- same bars as the benign encoder."""
    torch.cuda.set_device(0)
    sd = synth.outlier_encoder_state_dict()
    enc = ops.FrameEncoder(sd, chunk_frames=64)
    try:
        T, K, W = 48, 8, 3
        clip = synth.make_clip(11, T)
        frames = vit.image_processor_224(clip)
        oracle = vit.VitOracle(sd)
        torch.set_num_threads(os.cpu_count() or 1)
        hidden = oracle.forward_hidden(frames, post_ln=False)
        ratio = (hidden.abs().amax() / hidden.abs().median()).item()
        assert ratio >= 50.0, f"the stress weights must produce massive channels (max/median {ratio:.1f})"
        res = sas.sample_mdf_batch(clip.unsqueeze(0).cuda(), enc, K, W, want_aux=True)
        _, aux = mdf.sample_representative_frames(frames, oracle, K, W, {"Failure": 0, "Zeros": 0}, return_aux=True)
        cos = (res["feats"][0].cpu() * aux["feats"]).sum(dim=1)
        eps = (res["lcl_avg"][0].cpu() - aux["lcl_avg"]).abs().max().item()
        print(f"outlier encoder: max/median |x| {ratio:.0f}, min feature cosine {cos.min().item():.7f}, max |d lcl| {eps:.2e}")
        assert cos.min().item() >= 0.9999, cos.min().item()
        assert eps <= 1e-3, eps
        got, want = res["indices"][0].cpu().tolist(), list(aux["indices"])
        assert int(res["status"][0]) == aux["status"]
        assert not _excused(got, want, aux["lcl_avg"], eps), (got, want, eps)
        # the fp32 hidden state itself, layer by layer where the outliers enter (layers 2-5) and at the end
        patches = ops.preprocess_u8(clip[:8].cuda())
        for n_layers in (2, 6, 12):
            h_gpu = enc.hidden(patches, n_layers).cpu()
            h_ref = oracle.forward_hidden(frames[:8], n_layers=n_layers, post_ln=False)
            rel = ((h_gpu - h_ref).norm(dim=-1) / h_ref.norm(dim=-1)).max().item()
            print(f"  hidden state after {n_layers} blocks: max per-row relative error {rel:.3e}")
            assert rel <= 6e-2, (n_layers, rel)     # measured 3.6e-2 after block 6: bf16 operands under x50 rows of fc2
    finally:
        enc.close()


@pytest.mark.parametrize("K,W,want_status", [(16, 8, 1), (8, 2, 0)])
def test_all_identical_frames_clip(encoder, K, W, want_status):
    """Every interior score ties (utils.py:57-61 gives the same value for every window of identical features): whichever
    frames are picked, the stored frames are the same image, the status is fixed by (T, K, W), and every pick that differs
    from the oracle's is an exact-or-within-eps tie."""
    T = 64
    one = synth.make_clip(21, 1)
    clip = one.expand(T, -1, -1, -1).contiguous()
    frames = vit.image_processor_224(clip)
    res = sas.sample_mdf_batch(clip.unsqueeze(0).cuda(), encoder, K, W, want_aux=True)
    feats = res["feats"][0].cpu()
    assert torch.equal(feats, feats[:1].expand_as(feats)), "identical frames must give bit-identical features"
    _, aux = mdf.sample_representative_frames(frames, _FeatureModel(feats), K, W, {"Failure": 0, "Zeros": 0}, return_aux=True)
    assert int(res["status"][0]) == aux["status"] == want_status
    eps = (res["lcl_avg"][0].cpu() - aux["lcl_avg"]).abs().max().item()
    assert eps <= 1e-5
    got, want = res["indices"][0].cpu().tolist(), list(aux["indices"])
    assert len(set(got)) == K and all(0 <= g < T for g in got)
    assert not _excused(got, want, aux["lcl_avg"], max(eps, 1e-7)), (got, want, eps)
    assert torch.equal(res["frames"][0].cpu(), frames[:1].expand(K, -1, -1, -1)), "the stored frames are the one image"


def test_two_scene_clip_with_exact_ties(encoder):
    """Two scenes of identical frames: every interior score ties (inside AND across the scenes) and dips at the cut.  Picks
    may differ from the oracle only between tied frames; the spacing rule holds; no pick sits on the cut while tied interior
    frames are left."""
    T, K, W = 64, 6, 4
    a, b = synth.make_clip(31, 1), synth.make_clip(32, 1)
    clip = torch.cat([a.expand(T // 2, -1, -1, -1), b.expand(T // 2, -1, -1, -1)]).contiguous()
    frames = vit.image_processor_224(clip)
    res = sas.sample_mdf_batch(clip.unsqueeze(0).cuda(), encoder, K, W, want_aux=True)
    feats = res["feats"][0].cpu()
    _, aux = mdf.sample_representative_frames(frames, _FeatureModel(feats), K, W, {"Failure": 0, "Zeros": 0}, return_aux=True)
    assert int(res["status"][0]) == aux["status"] == 0
    eps = max((res["lcl_avg"][0].cpu() - aux["lcl_avg"]).abs().max().item(), 1e-7)
    got, want = res["indices"][0].cpu().tolist(), list(aux["indices"])
    assert not _excused(got, want, aux["lcl_avg"], eps), (got, want, eps)
    srt = sorted(got)
    assert all(y - x >= W for x, y in zip(srt, srt[1:])), got                      # utils.py:76-88 spacing
    lcl = aux["lcl_avg"]
    assert min(float(lcl[g]) for g in got) >= float(lcl[T // 2 - W + 1:T // 2 + W].min()), "a cut frame beat a tied interior frame"
    # (interior windows of BOTH scenes hold identical features, so their scores are all (2W - 1)/(2W - 1) = 1 up to rounding:
    # the tie spans the two scenes, and which scene wins is exactly the kind of difference the tolerance rule excuses)
    assert torch.equal(res["frames"][0].cpu(), frames[torch.tensor(got)])


@pytest.mark.parametrize("T,K,W", [(1, 1, 8), (2, 2, 0), (3, 2, 1), (3, 3, -1)])
def test_tiny_clips_end_to_end(encoder, T, K, W):
    clip = synth.make_clip(40 + T, T)
    frames = vit.image_processor_224(clip)
    res = sas.sample_mdf_batch(clip.unsqueeze(0).cuda(), encoder, K, W, want_aux=True)
    _, aux = mdf.sample_representative_frames(frames, _FeatureModel(res["feats"][0].cpu()), K, W, {"Failure": 0, "Zeros": 0},
                                              return_aux=True)
    assert int(res["status"][0]) == aux["status"]
    eps = max((res["lcl_avg"][0].cpu() - aux["lcl_avg"]).abs().max().item(), 1e-7)
    got, want = res["indices"][0].cpu().tolist(), list(aux["indices"])
    assert not _excused(got, want, aux["lcl_avg"], eps), (got, want)
    assert torch.equal(res["frames"][0].cpu(), frames[torch.tensor(got)])
    # and through the reference-signature wrapper: same frames, counters bumped where the reference bumps them
    dc = {"Failure": 0, "Zeros": 0}
    out = sas.sample_representative_frames(frames, encoder, K, W, dc)
    assert torch.equal(out, frames[torch.tensor(got)]) and dc["Failure"] == int(aux["status"] == 1)


@pytest.mark.parametrize("T,K,W", [(4097, 2048, 8), (10000, 2048, 1), (4097, 16, -1)])
def test_very_long_clips_and_k_2048(encoder, T, K, W):
    """> 4096 candidates: the fallback runs as K selection rounds instead of the shared-memory bitonic sort, the greedy
    kernel keeps up to K + 1 = 2049 open intervals.  Stage-wise parity: the restated sampler on the GPU's own features."""
    clip = synth.make_clip(50, T, device="cuda")                                   # 0.6 - 1.5 GB uint8, generated on the device
    res = sas.sample_mdf_batch(clip.unsqueeze(0), encoder, K, W, want_aux=True, want_frames=False)
    feats = res["feats"][0].cpu()
    dummy = torch.zeros(T, 1, 1, 1)
    _, aux = mdf.sample_representative_frames(dummy, _FeatureModel(feats), K, W, {"Failure": 0, "Zeros": 0}, return_aux=True)
    assert int(res["status"][0]) == aux["status"]
    eps = max((res["lcl_avg"][0].cpu() - aux["lcl_avg"]).abs().max().item(), 1e-7)
    assert eps <= 1e-5
    got, want = res["indices"][0].cpu().tolist(), list(aux["indices"])
    assert len(set(got)) == K
    bad = _excused(got, want, aux["lcl_avg"], eps)
    n_diff = sum(a != b for a, b in zip(got, want))
    print(f"T={T} K={K} W={W}: status {aux['status']}, {n_diff} of {K} picks differ (all within 2*eps = {2 * eps:.1e})")
    assert not bad, bad[:8]
    if aux["status"] == 1:       # plain top-K: a near-tie swaps two neighbours and nothing else
        assert n_diff <= max(2, K // 10), n_diff
    else:                        # greedy: a near-tie inside one interval changes the split and every later pick (971 of 2048
        srt = sorted(got)        # at T = 10 000, all within 2.6e-6) -- what must hold is the rule itself
        Wr = T // 20 if W == -1 else W
        assert all(y - x >= Wr for x, y in zip(srt, srt[1:])), "spacing rule violated"
    # stage-wise exactness (SURVEY 8(d) tolerance 1): the selection kernel on the ORACLE's fp32 scores picks the oracle's
    # indices, except between exactly equal scores
    idx, status = ops.mdf_select(aux["lcl_avg"].cuda(), K, T // 20 if W == -1 else W)
    assert int(status) == aux["status"]
    lcl = aux["lcl_avg"]
    exact = [(a, b) for a, b in zip(idx.cpu().tolist(), want) if a != b and float(lcl[a]) != float(lcl[b])]
    assert not exact, exact[:8]
    # the gather of those picks, straight from the uint8 clip
    sub = torch.tensor(got[:8], dtype=torch.int32, device="cuda").unsqueeze(0)
    fr = ops.gather_frames_u8(clip.unsqueeze(0), sub)[0].cpu()
    assert torch.equal(fr, vit.image_processor_224(clip[torch.tensor(got[:8], device="cuda")].cpu()))


def test_handle_serialises_across_streams(encoder):
    """sample_mdf_batch on a side stream with NO sync, then the host-buffer pipeline (the handle's own
    streams) on the same encoder -- both share one workspace and must come out as if run one after the other."""
    T, K, W = 64, 8, 4
    dev_clips = synth.make_clips(range(60, 68), T, device="cuda")
    host_clips = synth.make_clips(range(70, 78), T).pin_memory()
    want_dev = sas.sample_mdf_batch(dev_clips, encoder, K, W, want_aux=True)
    want_host = sas.sample_mdf_host(host_clips, encoder, K, W)
    torch.cuda.synchronize()
    want_dev = {k: v.clone() for k, v in want_dev.items() if v is not None}
    want_host = {k: v.clone() for k, v in want_host.items() if v is not None}
    side = torch.cuda.Stream()
    for _ in range(3):
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            got_dev = sas.sample_mdf_batch(dev_clips, encoder, K, W, want_aux=True)       # asynchronous, not synchronised
        got_host = sas.sample_mdf_host(host_clips, encoder, K, W)                          # returns after its own work
        with torch.cuda.stream(side):
            again = sas.sample_mdf_batch(dev_clips, encoder, K, W, want_aux=True)          # and back on the side stream
        torch.cuda.synchronize()
        for k in ("indices", "status", "frames"):
            assert torch.equal(got_host[k], want_host[k]), f"host pipeline raced: {k}"
        for k in ("indices", "status", "feats", "lcl_avg", "frames"):
            assert torch.equal(got_dev[k], want_dev[k]), f"device call raced: {k}"
            assert torch.equal(again[k], want_dev[k]), f"second device call raced: {k}"


def test_short_pick_rows_are_defined(encoder):
    """status 3 (T < K on the fallback path) leaves the unused tail of idx at -1, and the gather gives zero rows."""
    lcl = torch.rand(2, 5, device="cuda")
    idx, status = ops.mdf_select(lcl, 8, 8)
    assert status.tolist() == [3, 3]
    idx = idx.cpu()
    assert (idx[:, 1:] == -1).all() and (idx[:, 0] >= 0).all()
    clips = synth.make_clips(range(2), 5, device="cuda")
    fr = ops.gather_frames_u8(clips, idx.cuda())
    assert torch.count_nonzero(fr[:, 1:]) == 0
