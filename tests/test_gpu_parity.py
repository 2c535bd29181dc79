"""GPU parity tests (pytest -m gpu, run on the B200 box): every CUDA stage, called through the
C ABI, against the CPU oracle on the same seeded inputs and against the golden fixtures that
the reference itself produced (oracle/make_golden.py).  /root/reference is NOT needed here.

Tolerances:
  integer / index / byte stages ............ bit-exact (exact ties excused, see oracle/mdf.py)
  fp32 scores given identical fp32 features .. |d| <= 1e-5
  bf16 encoder vs fp32 reference ............. feature cosine >= 0.9999, Gram |d| <= 1e-3
  end-to-end indices ......................... identical except where the deciding scores differ
                                               by <= 2*eps, eps = max |lcl_gpu - lcl_ref| of that clip
"""
import os

import numpy as np
import pytest
import torch

from oracle import mdf, vit
import sasvqa_b200 as sas
from sasvqa_b200 import ops, sampler, synth

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def state_dict():
    return synth.random_encoder_state_dict(synth.REF_SEED)


@pytest.fixture(scope="module")
def encoder(state_dict):
    torch.cuda.set_device(0)
    enc = ops.FrameEncoder(state_dict, chunk_frames=96)      # small chunk: exercises multi-chunk + tail paths
    yield enc
    enc.close()


@pytest.fixture(scope="module")
def vit_oracle(state_dict):
    return vit.VitOracle(state_dict)


def _golden(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def assert_same_or_tied(got, want, scores, eps=0.0):
    assert len(got) == len(want), (got, want)
    excused = 0
    for a, b in zip(got, want):
        if a != b:
            assert abs(float(scores[a]) - float(scores[b])) <= eps, (got, want, float(scores[a]), float(scores[b]))
            excused += 1
    return excused


# ------------------------------------------------------------------------------------------ GEMM
def _gemm_reference(a, b, mode, vec, x0):
    acc = a.float() @ b.float().t()
    if mode == 0:
        return (acc + vec).to(torch.bfloat16)
    if mode == 1:
        y = acc + vec
        return (y * torch.sigmoid(1.702 * y)).to(torch.bfloat16)
    if mode == 2:
        return x0 + acc + vec
    M, N = acc.shape
    frames = M // 196
    out = x0.clone()
    out.view(frames, 197, N)[:, 1:, :] = acc.view(frames, 196, N) + vec[1:].unsqueeze(0)
    return out


@pytest.mark.parametrize("use_simt", [True, False], ids=["simt-check", "tcgen05"])
@pytest.mark.parametrize("M,N,K,mode", [
    (128, 256, 64, 0), (197, 768, 768, 0), (591, 2304, 768, 0), (1000, 3072, 768, 1),
    (4113, 768, 3072, 2), (197 * 5, 768, 768, 2), (196 * 3, 768, 768, 3), (196 * 40, 768, 768, 3),
    (128 * 149 + 5, 768, 768, 0),
])
def test_gemm_epilogues(M, N, K, mode, use_simt):
    g = torch.Generator(device=DEV).manual_seed(M * 7 + N + K + mode)
    a = (torch.randn(M, K, device=DEV, generator=g) * 0.5).to(torch.bfloat16)
    b = (torch.randn(N, K, device=DEV, generator=g) * 0.05).to(torch.bfloat16)
    if mode == 3:
        vec = torch.randn(197, N, device=DEV, generator=g)
        x0 = torch.randn((M // 196) * 197, N, device=DEV, generator=g)
    else:
        vec = torch.randn(N, device=DEV, generator=g)
        x0 = torch.randn(M, N, device=DEV, generator=g) if mode == 2 else None
    want = _gemm_reference(a, b, mode, vec, x0)
    out_f32 = x0.clone() if x0 is not None else None
    got = ops.test_gemm(a, b, mode, vec, out_f32=out_f32, use_simt=use_simt)
    torch.cuda.synchronize()
    tol = 2e-2 if mode in (0, 1) else 2e-3
    err = (got.float() - want.float()).abs().max().item()
    assert err <= tol * max(1.0, want.float().abs().max().item()), err
    if mode == 3:       # class-token rows untouched by the scatter
        assert torch.equal(got.view(-1, 197, N)[:, 0], x0.view(-1, 197, N)[:, 0])


def test_gemm_tcgen05_matches_check_kernel_closely():
    g = torch.Generator(device=DEV).manual_seed(5)
    a = torch.randn(777, 768, device=DEV, generator=g).to(torch.bfloat16)
    b = (torch.randn(2304, 768, device=DEV, generator=g) * 0.03).to(torch.bfloat16)
    bias = torch.randn(2304, device=DEV, generator=g)
    fast = ops.test_gemm(a, b, 0, bias, use_simt=False).float()
    slow = ops.test_gemm(a, b, 0, bias, use_simt=True).float()
    # same bf16 inputs, fp32 accumulate in both: only summation order and final bf16 rounding differ
    assert (fast - slow).abs().max().item() <= 2 ** -7 * max(1.0, slow.abs().max().item())


# ------------------------------------------------------------------------- LN / attention / K1 / K5
def test_layernorm_vs_torch():
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(1000, 768, device=DEV, generator=g) * 3 + 0.5
    gamma = torch.randn(768, device=DEV, generator=g)
    beta = torch.randn(768, device=DEV, generator=g)
    got = ops.test_layernorm(x, gamma, beta).float()
    want = torch.nn.functional.layer_norm(x, (768,), gamma, beta, 1e-5)
    assert (got - want).abs().max().item() <= 2 ** -8 * want.abs().max().item() + 1e-3


@pytest.mark.parametrize("impl,n", [(0, 3), (0, 40), (1, 3)], ids=["tcgen05", "tcgen05-persistent", "mma-check"])
def test_attention_vs_fp32_reference(impl, n):
    g = torch.Generator(device=DEV).manual_seed(2 + n)
    qkv = (torch.randn(n * 197, 2304, device=DEV, generator=g) * 1.5).to(torch.bfloat16)
    got = ops.test_attention(qkv, impl=impl).float().view(n, 197, 12, 64)
    q, k, v = [t.view(n, 197, 12, 64).transpose(1, 2) for t in qkv.float().split(768, dim=1)]
    att = torch.softmax((q @ k.transpose(-1, -2)) * 0.125, dim=-1)
    want = (att @ v).transpose(1, 2)
    assert (got - want).abs().max().item() <= 3e-2
    assert torch.nn.functional.cosine_similarity(got.reshape(n * 197, -1), want.reshape(n * 197, -1)).min() > 0.9995


def test_preprocess_u8_bit_exact_vs_oracle():
    u8 = synth.make_clip(3, 5)
    want = vit.image_processor_224(u8)                                        # fp32 CHW, CPU oracle
    patches = ops.preprocess_u8(u8.to(DEV)).cpu()
    ref = want.reshape(5, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(5 * 196, 768).to(torch.bfloat16)
    assert torch.equal(patches.view(torch.int16), ref.view(torch.int16))
    patches2 = ops.patchify_f32(want.to(DEV)).cpu()
    assert torch.equal(patches2.view(torch.int16), ref.view(torch.int16))
    # exhaustive: every one of the 3 x 256 possible channel values (the kernel's one-FMA form must round to the
    # same bf16 as the processor's multiply / subtract / divide)
    allv = (torch.arange(224 * 224 * 3, dtype=torch.int64) // 3 % 256).to(torch.uint8).reshape(1, 224, 224, 3)
    want = vit.image_processor_224(allv)
    ref = want.reshape(1, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(196, 768).to(torch.bfloat16)
    assert torch.equal(ops.preprocess_u8(allv.to(DEV)).cpu().view(torch.int16), ref.view(torch.int16))


def test_gather_bit_exact_vs_oracle():
    clips = torch.stack([synth.make_clip(10 + b, 6) for b in range(2)])          # [2, 6, 224, 224, 3]
    idx = torch.tensor([[5, 0, 3], [2, 2, 4]], dtype=torch.int32)
    got = ops.gather_frames_u8(clips.to(DEV), idx.to(DEV)).cpu()
    for b in range(2):
        want = vit.image_processor_224(clips[b])[idx[b].long()]
        assert torch.equal(got[b], want)
    f32 = torch.stack([vit.image_processor_224(clips[b]) for b in range(2)])
    got2 = ops.gather_frames_f32(f32.to(DEV), idx.to(DEV)).cpu()
    assert torch.equal(got2, got)


def test_preprocess_matches_hf_processor_fixture(golden_dir):
    g = _golden(golden_dir, "encoder_hf.npz")
    u8 = synth.make_clip(int(g["clip_id"]), int(g["T"]))
    idx = torch.arange(int(g["T"]), dtype=torch.int32).view(1, -1)
    px = ops.gather_frames_u8(u8.unsqueeze(0).to(DEV), idx.to(DEV))[0].cpu()
    assert np.abs(px[:, :, ::37, ::41].numpy() - g["pixel_probe"]).max() < 2e-6


# ------------------------------------------------------------------------------------ encoder
def test_encoder_hidden_states_vs_fp32_oracle(encoder, vit_oracle):
    u8 = synth.make_clip(7, 3)
    px = vit.image_processor_224(u8)
    patches = ops.preprocess_u8(u8.to(DEV))
    for n_layers, tol in ((0, 2e-2), (1, 4e-2), (12, 0.25)):
        got = encoder.hidden(patches, n_layers).cpu()
        want = vit_oracle.forward_hidden(px, n_layers=n_layers, post_ln=False)
        err = (got - want).abs().max().item()
        cos = torch.nn.functional.cosine_similarity(got.reshape(-1, 768), want.reshape(-1, 768)).min().item()
        assert err <= tol and cos > 0.9995, (n_layers, err, cos)


def test_encoder_features_vs_oracle_and_hf_fixture(encoder, vit_oracle, golden_dir):
    g = _golden(golden_dir, "encoder_hf.npz")
    u8 = synth.make_clip(int(g["clip_id"]), int(g["T"]))
    feats = encoder.features_u8(u8.to(DEV)).cpu()
    want = torch.from_numpy(g["feats"])                                         # HF GitVisionModel, fp32
    cos = (feats * want).sum(dim=1)
    assert cos.min().item() >= 0.9999, cos
    assert (feats.norm(dim=1) - 1).abs().max().item() < 1e-5
    assert ((feats @ feats.t()) - (want @ want.t())).abs().max().item() <= 1e-3
    oracle_feats = vit_oracle.features(vit.image_processor_224(u8))
    assert (oracle_feats * feats).sum(dim=1).min().item() >= 0.9999


def test_encoder_chunking_is_invisible(encoder):
    u8 = synth.make_clip(8, 200).to(DEV)                                        # 200 frames > chunk of 96
    full = encoder.features_u8(u8)
    part = torch.cat([encoder.features_u8(u8[:50]), encoder.features_u8(u8[50:])])
    assert torch.equal(full, part)


# ----------------------------------------------------------------------------- scores / selection
def test_scores_vs_oracle_fp32():
    rng = np.random.RandomState(3)
    for T, W in ((64, 8), (128, 8), (50, 4), (12, 0), (10, 8), (512, 8), (100, 5)):
        f = torch.nn.functional.normalize(torch.from_numpy(
            (np.cumsum(rng.randn(T, 768), axis=0) * 0.2 + rng.randn(1, 768)).astype(np.float32)))
        want = mdf.local_average(mdf.gram(f), W)
        lcl, gram = ops.mdf_scores(f.to(DEV), W, want_gram=True)
        assert (lcl.cpu() - want).abs().max().item() <= 1e-5
        assert (gram.cpu() - mdf.gram(f)).abs().max().item() <= 1e-5
        if W > 0:
            assert torch.all(lcl[:W] == 0) and torch.all(lcl[T - W:] == 0)      # exact-zero borders


def test_select_matches_reference_fixtures(golden_dir):
    g = _golden(golden_dir, "mdf_select.npz")
    D = int(g["D"])
    fo = io = 0
    n = n_fb = n_err = 0
    for T, K, W, failure, zeros, err, n_idx in g["meta"].tolist():
        feats = g["feats"][fo:fo + T * D].reshape(T, D)
        want = g["indices"][io:io + n_idx].tolist()
        fo += T * D
        io += n_idx
        if T == 0:
            continue
        f = torch.nn.functional.normalize(torch.from_numpy(feats.copy()))
        Wr = mdf.resolve_window(W, T)
        lcl = mdf.local_average(mdf.gram(f), Wr)                                # identical fp32 scores on both sides
        idx, status = ops.mdf_select(lcl.to(DEV), K, Wr)
        n += 1
        if err:
            assert int(status) == ops.STATUS_TOO_FEW
            n_err += 1
            continue
        assert int(status) == (ops.STATUS_FALLBACK if failure else ops.STATUS_OK)
        n_fb += failure
        assert_same_or_tied(idx.cpu().tolist(), want, lcl)                      # bit-exact; exact ties excused
    assert n > 150 and n_fb > 10 and n_err > 5


def test_select_batched_equals_oracle_random():
    rng = np.random.RandomState(11)
    for T, K, W in ((128, 16, 8), (128, 8, 8), (512, 32, 8), (64, 16, 8), (300, 10, 3), (5000, 64, 16)):
        lcl = torch.from_numpy(rng.rand(9, T).astype(np.float32))
        if W:
            lcl[:, :W] = 0
            lcl[:, T - W:] = 0
        idx, status = ops.mdf_select(lcl.to(DEV), K, W)
        for b in range(9):
            want, st = mdf.mdf_select(lcl[b], K, W)
            assert int(status[b]) == st
            assert idx[b].cpu().tolist() == want                               # continuous scores: no ties at all


def test_topk_strided_fixture_and_large(golden_dir):
    g = _golden(golden_dir, "samplers_misc.npz")
    o = so = 0
    for T, K, ds in g["mif_meta"].tolist():
        s = torch.from_numpy(g["mif_scores"][so:so + T].copy())
        assert sas.mif_select(s.to(DEV), K, ds) == g["mif_idx"][o:o + K].tolist()
        o += K
        so += T
    rng = np.random.RandomState(12)
    for T, K, ds in ((10000, 50, 1), (10000, 7, 3), (4096, 4096, 1), (4097, 100, 1)):   # rounds kernel / full sort
        s = torch.from_numpy(rng.randn(3, T).astype(np.float32))
        got = sas.mif_select(s.to(DEV), K, ds).cpu()
        for b in range(3):
            assert got[b].tolist() == mdf.mif_select(s[b].numpy(), K, ds)
    ties = torch.tensor([[0.0, -0.0, 1.0, 1.0, float("-inf"), 0.5, 1.0, 0.0]])
    assert sas.mif_select(ties.to(DEV), 8, 1).cpu()[0].tolist() == mdf.mif_select(ties[0].numpy(), 8, 1)


# ------------------------------------------------------------------------------- end to end
def test_e2e_vs_reference_function_fixture(encoder, golden_dir):
    """Our sample_representative_frames vs the reference's own function driven with HF's fp32
    encoder (fixture written by oracle/make_golden.py)."""
    g = _golden(golden_dir, "mdf_e2e_hf.npz")
    excused_total = 0
    for tag in ("c1", "t64k8w4", "t128k8w8"):
        cid, T, K, W, failure = g[tag + "_meta"].tolist()
        u8 = synth.make_clip(cid, T)
        frames = vit.image_processor_224(u8)                                    # what the reference was fed
        res = ops.mdf_sample_device(encoder, frames.unsqueeze(0).to(DEV), K, W, want_frames=True, want_aux=True)
        lcl_gpu = res["lcl_avg"][0].cpu()
        lcl_ref = torch.from_numpy(g[tag + "_lcl"])
        eps = (lcl_gpu - lcl_ref).abs().max().item()
        assert eps <= 1e-3, eps
        feats_ref = torch.from_numpy(g[tag + "_feats"])
        assert (res["feats"][0].cpu() * feats_ref).sum(dim=1).min().item() >= 0.9999
        assert int(res["status"][0]) == failure
        got = res["indices"][0].cpu().tolist()
        excused_total += assert_same_or_tied(got, g[tag + "_indices"].tolist(), lcl_ref, eps=2 * eps)
        # returned frames are the reference's frames[res], bit for bit
        assert torch.equal(res["frames"][0].cpu(), frames[torch.tensor(got)])
        # and the drop-in signature gives the same thing, with the reference's counter semantics
        dc = {"Failure": 0, "Zeros": 0}
        out = sampler.sample_representative_frames(frames, encoder, K, W, dc)
        assert out.device == frames.device and torch.equal(out, frames[torch.tensor(got)])
        assert dc["Failure"] == failure
    print("tolerance-excused index mismatches:", excused_total)


def test_e2e_u8_batch_host_and_oracle_agree(encoder, vit_oracle):
    T, K, W = 48, 6, 3
    clips = torch.stack([synth.make_clip(20 + b, T) for b in range(3)])
    dev = sas.sample_mdf_batch(clips.to(DEV), encoder, K, W, want_aux=True)
    host = sas.sample_mdf_host(clips.pin_memory(), encoder, K, W)
    assert torch.equal(dev["indices"].cpu(), host["indices"])
    assert torch.equal(dev["status"].cpu(), host["status"])
    assert torch.equal(dev["frames"].cpu(), host["frames"])
    for b in range(3):
        frames = vit.image_processor_224(clips[b])
        _, aux = mdf.sample_representative_frames(frames, vit_oracle, K, W, {"Failure": 0, "Zeros": 0}, return_aux=True)
        eps = (dev["lcl_avg"][b].cpu() - aux["lcl_avg"]).abs().max().item()
        assert eps <= 1e-3
        assert_same_or_tied(dev["indices"][b].cpu().tolist(), aux["indices"], aux["lcl_avg"], eps=2 * eps)
        assert int(dev["status"][b]) == aux["status"]


def test_edge_cases_through_public_api(encoder):
    dc = {"Failure": 0, "Zeros": 0}
    # empty clip list / empty clip
    res = sas.sample_mdf_batch(torch.zeros(2, 0, 224, 224, 3, dtype=torch.uint8, device=DEV), encoder, 4, 8, dc)
    assert res["status"].cpu().tolist() == [2, 2] and float(res["frames"].abs().sum()) == 0 and dc["Zeros"] == 2
    # T < K on the fallback path: the reference raises
    frames = vit.image_processor_224(synth.make_clip(30, 5))
    with pytest.raises(RuntimeError):
        sas.sample_representative_frames(frames, encoder, 16, 8, dc)
    # adaptive width W = -1 with T < 20 -> W = 0 -> all scores 1.0
    res = sas.sample_mdf_batch(synth.make_clip(31, 12).unsqueeze(0).to(DEV), encoder, 4, -1, want_aux=True)
    assert torch.all(res["lcl_avg"] == 1.0)
    # wrong frame size: same complaint as HF's embedding layer
    with pytest.raises(ValueError):
        sas.sample_representative_frames(torch.zeros(4, 3, 128, 128), encoder, 2, 1, dc)


# ------------------------------------------------------- full-size, size-independent properties
def test_full_size_properties_c2_like(encoder):
    """BASELINE config 2 shape per clip (T=128, K=16) over a batch too large for the CPU oracle:
    checked through invariants -- index range/uniqueness, spacing rule on greedy clips,
    top-K property on fallback clips, batch-split invariance, gathered frames == source frames."""
    B, T, K = 24, 128, 16
    clips = synth.make_clips(range(100, 100 + B), T, device=DEV)
    for W in (8, 4):
        res = sas.sample_mdf_batch(clips, encoder, K, W, want_aux=True)
        idx, st, lcl = res["indices"].cpu(), res["status"].cpu(), res["lcl_avg"].cpu()
        assert idx.min() >= 0 and idx.max() < T
        for b in range(B):
            picks = idx[b].tolist()
            assert len(set(picks)) == K
            if st[b] == 0:
                assert min(abs(p - q) for i, p in enumerate(picks) for q in picks[:i]) >= W
                assert picks[0] == int(lcl[b].argmax())
            else:
                assert st[b] == 1
                kth = lcl[b][picks].min()
                assert (lcl[b] > kth).sum() < K and torch.all(lcl[b][picks][:-1] >= lcl[b][picks][1:])
        halves = [sas.sample_mdf_batch(clips[s], encoder, K, W) for s in (slice(0, 7), slice(7, B))]
        assert torch.equal(torch.cat([h["indices"] for h in halves]).cpu(), idx)
        want = synth.normalize_frames_reference(clips[3][idx[3].to(DEV).long()])
        assert (res["frames"][3] - want).abs().max().item() <= 2e-6


# ------------------------------------------------------------------ K0: resize + centre crop (row f3)
def test_resize_crop_bit_exact_vs_oracle_and_hf_fixture(golden_dir):
    from oracle import resize
    g = _golden(golden_dir, "resize_hf.npz")
    for h, w in resize.RESIZE_CASES:
        frames = resize.resize_case_frames(h, w)
        got = ops.resize_crop_u8(torch.from_numpy(frames).to(DEV)).cpu().numpy()
        assert np.array_equal(got, g[f"out_{h}x{w}"]), (h, w)             # == the HF image processor's own output
    # sizes the fixture does not hold (large down-scale, portrait, tiny up-scale, odd sizes; several frames)
    for h, w, n in [(720, 1280, 2), (640, 360, 3), (64, 48, 2), (251, 333, 5), (1080, 1920, 1)]:
        rng = np.random.RandomState(h * 3 + w)
        frames = rng.randint(0, 256, (n, h, w, 3), dtype=np.uint8)
        got = ops.resize_crop_u8(torch.from_numpy(frames).to(DEV)).cpu().numpy()
        assert np.array_equal(got, resize.resize_crop_u8(frames)), (h, w)
    same = torch.randint(0, 256, (3, 224, 224, 3), dtype=torch.uint8, device=DEV)
    assert torch.equal(ops.resize_crop_u8(same), same)                    # 224x224: resize and crop are identities


def test_e2e_non_224_clips_equal_processed_clips(encoder):
    """Decoded frames of another size go through K0 -> the result must be exactly what the path gives for the
    clip the host image processor would have produced (oracle resize), on the device and the host entry points."""
    from oracle import resize
    T, K, W = 40, 5, 3
    for h, w in [(240, 320), (300, 226)]:
        raw = torch.stack([synth.make_clip(70 + b, T, H=h, W=w) for b in range(2)])            # [2, T, h, w, 3]
        processed = torch.from_numpy(resize.resize_crop_u8(raw.numpy()))                          # [2, T, 224, 224, 3]
        want = sas.sample_mdf_batch(processed.to(DEV), encoder, K, W, want_aux=True)
        got = sas.sample_mdf_batch(raw.to(DEV), encoder, K, W, want_aux=True)
        host = sas.sample_mdf_host(raw.pin_memory(), encoder, K, W)
        for key in ("indices", "status", "frames", "feats", "lcl_avg"):
            assert torch.equal(got[key], want[key]), (key, h, w)
        for key in ("indices", "status", "frames"):
            assert torch.equal(host[key], want[key].cpu()), (key, h, w)
        assert torch.equal(got["frames"][1].cpu(),
                           synth.normalize_frames_reference(processed[1][got["indices"][1].cpu().long()]))


# ------------------------------------------------------------------ MIF, BASELINE config 3 (embedding relevance)
def test_mif_scores_and_selection_vs_oracle():
    g = torch.Generator().manual_seed(5)
    B, T = 6, 128
    feats = torch.nn.functional.normalize(torch.randn(B, T, 768, generator=g), dim=-1)
    q = synth.question_embeddings(range(B))
    got = ops.mif_scores(feats.to(DEV), q.to(DEV)).cpu()
    for b in range(B):
        want = mdf.mif_scores(feats[b], q[b])
        assert (got[b] - want).abs().max().item() <= 1e-5
        for K, ds in [(8, 1), (8, 2), (5, 3), (1, 1)]:                       # selection on the GPU's own scores: exact
            idx = ops.topk_strided(got[b:b + 1].to(DEV), K, ds)[0].cpu().tolist()
            assert idx == mdf.mif_select(got[b].numpy(), K, ds)


def test_mif_e2e_c3_like(encoder, vit_oracle):
    """clips -> encoder -> <feat, q> -> strided top-K, against the fp32 oracle chain on the same clips."""
    T, K = 32, 8
    ids = [40, 41]
    clips = torch.stack([synth.make_clip(c, T) for c in ids])
    q = synth.question_embeddings(ids)
    for ds in (1, 2):
        res = sas.sample_mif_batch(clips.to(DEV), encoder, q, K, ds, want_frames=True, want_aux=True)
        for b in range(len(ids)):
            frames = vit.image_processor_224(clips[b])
            feats = vit_oracle.features(frames)
            want_scores = mdf.mif_scores(feats, q[b])
            got_scores = res["scores"][b].cpu()
            eps = (got_scores - want_scores).abs().max().item()
            assert eps <= 1e-3                                                # bf16 encoder vs fp32, as for the Gram
            got = res["indices"][b].cpu().tolist()
            assert got == mdf.mif_select(got_scores.numpy(), K, ds)           # exact on its own scores
            assert_same_or_tied(got, mdf.mif_select(want_scores.numpy(), K, ds), want_scores, eps=2 * eps)
            assert all(i % ds == 0 for i in got)
            assert torch.equal(res["frames"][b].cpu(), frames[torch.tensor(got)])
    with pytest.raises(sas.SasvqaError):
        sas.sample_mif_batch(clips.to(DEV), encoder, q, 20, 2)                # K > ceil(T / ds_rate): topk raises


def test_full_size_properties_c4_long_video(encoder):
    """BASELINE config 4 shape (T=512, K=32, W=8): spacing rule / first pick / chunk invariance on long clips."""
    B, T, K, W = 2, 512, 32, 8
    clips = synth.make_clips(range(300, 300 + B), T, device=DEV)
    res = sas.sample_mdf_batch(clips, encoder, K, W, want_aux=True)
    idx, st, lcl = res["indices"].cpu(), res["status"].cpu(), res["lcl_avg"].cpu()
    for b in range(B):
        picks = idx[b].tolist()
        assert len(set(picks)) == K and min(picks) >= 0 and max(picks) < T
        assert torch.all(lcl[b][:W] == 0) and torch.all(lcl[b][T - W:] == 0)
        if st[b] == 0:
            assert picks[0] == int(lcl[b].argmax())
            assert min(abs(p - r) for i, p in enumerate(picks) for r in picks[:i]) >= W
        want, status = mdf.mdf_select(lcl[b], K, W)                         # oracle selection on the GPU's scores
        assert status == int(st[b])
        assert_same_or_tied(picks, want, lcl[b])
    one = sas.sample_mdf_batch(clips[1:2], encoder, K, W)
    assert torch.equal(one["indices"].cpu(), idx[1:2])


# ------------------------------------------------------------------ row f2: visual tokens of the sampled frames
def test_visual_tokens_vs_fp32_oracle_and_hf_fixture(encoder, vit_oracle, golden_dir):
    """bf16 encoder + projection vs the fp32 chain of src/modeling/modeling.py:76-95.  Tolerance: the outputs are
    LayerNorm-ed (unit scale); bf16 activations give |d| <= 0.06 per element, per-token cosine >= 0.9995."""
    g = _golden(golden_dir, "visual_tokens_hf.npz")
    psd = synth.random_projection_state_dict()
    encoder.set_projection(*[psd[f"visual_projection.{k}"] for k in ("0.weight", "0.bias", "1.weight", "1.bias")])
    clips = torch.stack([synth.make_clip(int(c), int(g["T"])) for c in g["clip_ids"]])          # [2, 2, 224, 224, 3] uint8
    frames = torch.stack([vit.image_processor_224(c) for c in clips])                             # [2, 2, 3, 224, 224]
    for project, probe in ((False, g["hidden_probe"]), (True, g["tokens_probe"])):
        want = vit.visual_tokens(frames, vit_oracle, psd if project else None)
        got = sas.encode_sampled_frames(frames.to(DEV), encoder, project=project).cpu()
        assert got.shape == want.shape == (2, 2 * 197, 768)
        cos = torch.nn.functional.cosine_similarity(got, want, dim=-1).min().item()
        err = (got - want).abs().max().item()
        assert cos >= 0.9995 and err <= 0.06, (project, cos, err)
        assert np.abs(got[:, ::29, ::48].numpy() - probe).max() <= 0.06                         # HF's own output
        got_u8 = encoder.visual_tokens(clips.to(DEV), project=project).reshape(2, 2 * 197, 768).cpu()
        assert torch.equal(got_u8, got)                                                           # uint8 entry == fp32 entry
    with pytest.raises(ValueError):
        sas.encode_sampled_frames(frames[0].to(DEV), encoder)


# ------------------------------------------------------------------ extraction loop + on-disk artefacts
def test_generate_h5_rows_match_reference_style_loop(encoder, tmp_path):
    """generate_h5 (extract_features.py:41-111 over decoded clips) == one reference-signature sampler call per
    clip, rows stored as extract_features.py:96-97 does; mixed frame sizes, an empty clip, all three strategies."""
    from sasvqa_b200 import writer
    K, W = 4, 2
    clips = [synth.make_clip(80, 24), synth.make_clip(81, 24), synth.make_clip(82, 20, H=240, W=320),
             torch.zeros(0, 224, 224, 3, dtype=torch.uint8), synth.make_clip(83, 24), synth.make_clip(84, 31),
             synth.make_clip(85, 12, H=240, W=320), synth.make_clip(86, 19, H=240, W=320)]     # ragged groups of both sizes
    path = str(tmp_path / "msvd_qa_video_feat.h5")
    res = sas.generate_h5(clips, encoder, K, W, path, inds_outfile=str(tmp_path / "mdf_inds.json"))
    ds = writer.open_sampled_frames(path)
    assert ds.shape == (len(clips), K, 3 * 224 * 224) and res["debug_counter"]["Zeros"] == 1
    from oracle import resize
    dc = {"Failure": 0, "Zeros": 0}
    for i, clip in enumerate(clips):
        u8 = torch.from_numpy(resize.resize_crop_u8(clip.numpy())) if clip.shape[0] else clip
        frames = vit.image_processor_224(u8) if clip.shape[0] else torch.zeros(0, 3, 224, 224)
        want = sas.sample_representative_frames(frames, encoder, K, W, dc)          # the reference call, per clip
        assert np.array_equal(np.asarray(ds[i]), want.reshape(K, -1).numpy()), i
    assert dc == res["debug_counter"]
    import json
    assert json.load(open(tmp_path / "mdf_inds.json"))["1"] == res["indices"][1].tolist()
    for strategy in ("uni", "git6"):
        np.random.seed(666)
        got = sas.generate_h5(clips[:2], None, K, W, str(tmp_path / f"{strategy}.h5"), sampling_strategy=strategy)
        ds2 = writer.open_sampled_frames(str(tmp_path / f"{strategy}.h5"))
        np.random.seed(666)
        for i in range(2):
            frames = vit.image_processor_224(clips[i])
            want = sas.sample_frames_uniform(frames, K) if strategy == "uni" else \
                sas.sample_frame_indices(frames, K, 4, len(frames))
            assert np.array_equal(np.asarray(ds2[i]), want.reshape(K, -1).numpy()), (strategy, i)


@pytest.mark.parametrize("W", [3, -1], ids=["W3", "adaptive"])
def test_ragged_batch_equals_one_call_per_clip(encoder, W):
    """Clips of different lengths in ONE library call == the reference's one-video-per-call loop
    (extract_features.py:80-97), bit for bit: an empty clip, T < K (fallback raises there: TOO_FEW), T <= 2W (all-zero
    scores), a fallback clip, long clips; per-clip adaptive window with W = -1."""
    K = 6
    lengths = [40, 0, 17, 64, 5, 23, 128, 9, 6]
    clips = [synth.make_clip(300 + i, T) for i, T in enumerate(lengths)]
    dc = {"Failure": 0, "Zeros": 0}
    res = sas.sample_mdf_ragged(clips, encoder, K, W, debug_counter=dc, want_aux=True)
    assert res["offsets"].tolist() == [0] + list(np.cumsum(lengths))
    n_fail = 0
    for b, (clip, T) in enumerate(zip(clips, lengths)):
        lo, hi = int(res["offsets"][b]), int(res["offsets"][b + 1])
        if T == 0:
            assert int(res["status"][b]) == ops.STATUS_EMPTY and res["indices"][b].cpu().tolist() == [-1] * K
            assert float(res["frames"][b].abs().sum()) == 0.0
            continue
        one = sas.sample_mdf_batch(clip.unsqueeze(0).to(DEV), encoder, K, W, want_aux=True)
        assert int(res["status"][b]) == int(one["status"][0]), (b, T)
        assert torch.equal(res["feats"][lo:hi], one["feats"][0]) and torch.equal(res["lcl_avg"][lo:hi], one["lcl_avg"][0])
        if int(one["status"][0]) != ops.STATUS_TOO_FEW:
            assert torch.equal(res["indices"][b], one["indices"][0]), (b, T)
            assert torch.equal(res["frames"][b], one["frames"][0])
        n_fail += int(one["status"][0]) == ops.STATUS_FALLBACK
    assert dc == {"Failure": n_fail, "Zeros": 1}
    # CPU clips without aux outputs take the host-buffer pipeline (groups of whole clips up to chunk_frames = 96 frames:
    # [40, 0, 17] [64, 5, 23] [128] [9, 6]); same answers, in host tensors
    host = sas.sample_mdf_ragged(clips, encoder, K, W)
    assert not host["indices"].is_cuda and torch.equal(host["status"], res["status"].cpu())
    ok = (res["status"] != ops.STATUS_TOO_FEW).cpu()
    assert torch.equal(host["indices"][ok], res["indices"].cpu()[ok]) and torch.equal(host["frames"][ok], res["frames"].cpu()[ok])
    empty = sas.sample_mdf_ragged([torch.zeros(0, 224, 224, 3, dtype=torch.uint8)] * 2, encoder, K, W)
    assert empty["status"].tolist() == [ops.STATUS_EMPTY] * 2 and float(empty["frames"].abs().sum()) == 0.0
    if W == 3:                              # T = 5 < K = 6 on the fallback path: the reference's topk raises
        assert int(res["status"][4]) == ops.STATUS_TOO_FEW
    else:                                   # T // 20 = 0 for the short clips: a zero-width window re-offers the first
        assert int(res["status"][4]) == ops.STATUS_OK      # maximum for ever (utils.py:67-88), K picks of one frame
        assert len(set(res["indices"][4].cpu().tolist())) == 1
    # decoded frames of another size: K0 on the chunks and on the picks, offsets carried through the pick map
    small = [synth.make_clip(320 + i, T, H=120, W=160) for i, T in enumerate((14, 33, 8))]
    rs = sas.sample_mdf_ragged(small, encoder, 4, 2)
    for b, clip in enumerate(small):
        one = sas.sample_mdf_batch(clip.unsqueeze(0).to(DEV), encoder, 4, 2)
        assert torch.equal(rs["indices"][b].cpu(), one["indices"][0].cpu()) and torch.equal(rs["frames"][b].cpu(), one["frames"][0].cpu())
    rd = sas.sample_mdf_ragged([c.to(DEV) for c in small], encoder, 4, 2)         # GPU clips: the device entry point
    assert rd["indices"].is_cuda and torch.equal(rd["indices"].cpu(), rs["indices"]) and torch.equal(rd["frames"].cpu(), rs["frames"])
    with pytest.raises(ValueError):
        sas.sample_mdf_ragged([clips[0], small[0]], encoder, K, W)                   # two frame sizes in one batch


def test_plain_c_client_on_the_gpu(abi_check_exe):
    """tests/c_abi/abi_check.c --gpu: selection entry points driven from C99 with cudaMalloc'd buffers."""
    import subprocess
    out = subprocess.run([abi_check_exe, "--gpu"], capture_output=True, text=True)
    assert out.returncode == 0 and "abi_check ok (gpu)" in out.stdout, out.stdout + out.stderr
