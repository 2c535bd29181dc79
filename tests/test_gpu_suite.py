"""The synthetic suite of BASELINE.json's north star ("identical frame indices on the synthetic suite"): whole clips
through the GPU path and through the CPU oracle (fp32 encoder restatement pinned to HF + the restated sampler pinned
to the reference's own function), compared pick by pick.  pytest -m gpu; the default is 16 scene-structured 128-frame
clips plus 2 clips of 512 frames (BASELINE config 4, K = 32); larger runs (SASVQA_SUITE_CLIPS / SASVQA_SUITE_LONG_CLIPS) are
recorded under profiles/ (summary JSON written to gpurun_out/).

Every clip is sampled under four (K, W) settings that share its features: (16, 8) = BASELINE config 2 (always the
top-K fallback), (16, 4) and (8, 8) (greedy path), (16, -1) (adaptive window).  A mismatch is excused only if the
oracle's own scores of the two picks differ by <= 2 * eps, eps = max |lcl_gpu - lcl_oracle| of that clip.
"""
import json
import os

import pytest
import torch

from oracle import mdf, vit
import sasvqa_b200 as sas
from sasvqa_b200 import ops, synth

pytestmark = pytest.mark.gpu

SETTINGS = [(16, 8), (16, 4), (8, 8), (16, -1)]


class _CachedFeatures:
    """Oracle model stand-in that returns the fp32 last_hidden_state computed once per clip, chunk by chunk in the order
    the reference's loop asks for it (utils.py:37-40: serial chunks of 256 frames)."""

    def __init__(self, hidden):
        self.hidden = hidden
        self.cursor = 0

    def __call__(self, frames):
        n = frames.shape[0]
        out = self.hidden[self.cursor:self.cursor + n]
        self.cursor = (self.cursor + n) % self.hidden.shape[0]
        assert out.shape[0] == n
        return type("O", (), {"last_hidden_state": out})()


def _run_suite(n_clips, T, settings, enc, oracle_cpu, oracle_gpu, summary, cpu_checked_clips):
    """n_clips scene-structured clips of T frames through the GPU path and the fp32 oracle.  The oracle's encoder runs on
    the host for the first `cpu_checked_clips` clips (the pinned CPU restatement) and, for all clips, as the SAME
    plain-torch fp32 code on the device with TF32 off -- the two are compared where both exist, which is what allows a
    suite of this size inside the test budget (the host needs ~4 s per 128-frame clip)."""
    for c in range(n_clips):
        clip = synth.make_clip(1000 + c, T)
        frames = vit.image_processor_224(clip)
        gpu_clip = clip.unsqueeze(0).cuda()
        with torch.no_grad():
            fr_dev = vit.image_processor_224(gpu_clip[0])
            hidden = torch.cat([oracle_gpu(fr_dev[i:i + 64]).last_hidden_state for i in range(0, T, 64)]).cpu()
            if c < cpu_checked_clips:
                hidden_cpu = torch.cat([oracle_cpu(frames[i:i + 32]).last_hidden_state for i in range(0, T, 32)])
                f_a = torch.nn.functional.normalize(hidden.mean(dim=1))
                f_b = torch.nn.functional.normalize(hidden_cpu.mean(dim=1))
                drift = (f_a - f_b).abs().max().item()
                summary["max_oracle_gpu_vs_cpu_feature_diff"] = max(summary["max_oracle_gpu_vs_cpu_feature_diff"], drift)
                assert drift <= 2e-5, f"fp32 oracle on the device drifted from the CPU oracle: {drift}"
                hidden = hidden_cpu                                   # where the CPU oracle exists it IS the judge
        for K, W in settings:
            res = sas.sample_mdf_batch(gpu_clip, enc, K, W, want_aux=True)
            _, aux = mdf.sample_representative_frames(frames, _CachedFeatures(hidden), K, W, {"Failure": 0, "Zeros": 0},
                                                      return_aux=True)
            lcl_ref = aux["lcl_avg"]
            eps = (res["lcl_avg"][0].cpu() - lcl_ref).abs().max().item()
            cos = (res["feats"][0].cpu() * aux["feats"]).sum(dim=1).min().item()
            got, want = res["indices"][0].cpu().tolist(), list(aux["indices"])
            assert eps <= 1e-3 and cos >= 0.9999, (c, K, W, eps, cos)
            key = f"T{T}_K{K}_W{W}"
            st = summary["per_setting"].setdefault(key, {"picks": 0, "identical": 0, "excused": 0, "fallback_clips": 0,
                                                         "fallback_zero_border_clips": 0})
            st["fallback_clips"] += int(aux["status"] == 1)
            Wr = T // 20 if W == -1 else W
            st["fallback_zero_border_clips"] += int(aux["status"] == 1 and any(g < Wr or g >= T - Wr for g in got))
            summary["status_mismatches"] += int(int(res["status"][0]) != aux["status"])
            for a, b in zip(got, want):
                st["picks"] += 1
                summary["picks"] += 1
                if a == b:
                    st["identical"] += 1
                    summary["identical"] += 1
                else:
                    assert abs(float(lcl_ref[a]) - float(lcl_ref[b])) <= 2 * eps, (c, K, W, got, want, eps)
                    st["excused"] += 1
                    summary["excused"] += 1
            summary["max_eps"] = max(summary["max_eps"], eps)
            summary["min_feature_cosine"] = min(summary["min_feature_cosine"], cos)


def test_synthetic_suite_indices_identical():
    """Default (what the driver runs): 16 clips x 128 frames under four (K, W) settings + 2 clips x 512 frames (BASELINE
    config 4) under three = 16*56 + 2*80 = 1056 picks.  SASVQA_SUITE_CLIPS / SASVQA_SUITE_LONG_CLIPS scale it."""
    n_clips = int(os.environ.get("SASVQA_SUITE_CLIPS", "16"))
    n_long = int(os.environ.get("SASVQA_SUITE_LONG_CLIPS", "2"))
    torch.cuda.set_device(0)
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False                      # the device-side oracle is strict fp32
    torch.backends.cudnn.allow_tf32 = False
    sd = synth.random_encoder_state_dict(synth.REF_SEED)
    enc = ops.FrameEncoder(sd, chunk_frames=512)
    oracle_cpu, oracle_gpu = vit.VitOracle(sd), vit.VitOracle(sd, device="cuda")
    torch.set_num_threads(os.cpu_count() or 1)
    summary = {"clips_T128": n_clips, "clips_T512": n_long, "picks": 0, "identical": 0, "excused": 0, "max_eps": 0.0,
               "min_feature_cosine": 1.0, "status_mismatches": 0, "max_oracle_gpu_vs_cpu_feature_diff": 0.0, "per_setting": {}}
    try:
        _run_suite(n_clips, 128, SETTINGS, enc, oracle_cpu, oracle_gpu, summary, cpu_checked_clips=2)
        _run_suite(n_long, 512, [(32, 8), (16, 8), (32, -1)], enc, oracle_cpu, oracle_gpu, summary, cpu_checked_clips=1)
    finally:
        enc.close()
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    summary["excused_fraction"] = summary["excused"] / max(summary["picks"], 1)
    print("synthetic suite:", json.dumps(summary))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, f"parity_suite_{n_clips}x128_{n_long}x512.json"), "w") as f:
            json.dump(summary, f, indent=1)
    assert summary["status_mismatches"] == 0
    # ties within tolerance stay rare: measured 0.7 % at T=128 and 2.5 % at T=512 in round 1 -> 1.5 % over this mix
    assert summary["excused"] <= 0.015 * summary["picks"], summary


def test_full_size_c2_batch_indices_vs_oracle():
    """BASELINE configs[1] at its FULL size -- 256 clips x 128 frames, K = 16, W = 8, one library call -- against the oracle,
    pick by pick (this shape was only checked through invariants).  The oracle's fp32 encoder runs as the
    same plain-torch restatement on the device (TF32 off; agreement with the CPU restatement is asserted on two clips), the
    restated sampler (oracle/mdf.py == utils.py:31-94) on the host.  SASVQA_FULL_SHAPE = c3 | c4 runs the other full-size
    BASELINE configs the same way (c3: MIF on synthetic question embeddings, K = 8; c4: 64 clips x 512 frames, K = 32) --
    recorded under profiles/, not part of the default run; SASVQA_FULL_C2_CLIPS shrinks the clip count for quick runs."""
    shape = os.environ.get("SASVQA_FULL_SHAPE", "c2")
    n_default, T, K, W = {"c2": (256, 128, 16, 8), "c3": (256, 128, 8, 0), "c4": (64, 512, 32, 8)}[shape]
    n_clips = int(os.environ.get("SASVQA_FULL_C2_CLIPS", str(n_default)))
    torch.cuda.set_device(0)
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sd = synth.random_encoder_state_dict(synth.REF_SEED)
    enc = ops.FrameEncoder(sd, chunk_frames=2048)
    oracle_cpu, oracle_gpu = vit.VitOracle(sd), vit.VitOracle(sd, device="cuda")
    torch.set_num_threads(os.cpu_count() or 1)
    try:
        clips = synth.make_clips(range(n_clips), T, device="cuda")                  # the bench's own clip ids
        if shape == "c3":
            q = synth.question_embeddings(range(n_clips), device="cuda")
            res = sas.sample_mif_batch(clips, enc, q, K, 1, want_aux=True)
            got_lcl, got_st = res["scores"].cpu(), torch.zeros(n_clips, dtype=torch.int32)
        else:
            res = sas.sample_mdf_batch(clips, enc, K, W, want_aux=True)
            got_lcl, got_st = res["lcl_avg"].cpu(), res["status"].cpu()
        got_idx, got_feats = res["indices"].cpu(), res["feats"].cpu()
        picks = identical = excused = status_bad = cascade = clips_with_tie = 0
        max_eps, min_cos = 0.0, 1.0
        for c in range(n_clips):
            with torch.no_grad():
                fr = vit.image_processor_224(clips[c])
                pooled = torch.cat([oracle_gpu(fr[i:i + 64]).last_hidden_state.mean(dim=1) for i in range(0, T, 64)]).cpu()
                if c < 2 and T <= 128:                                              # device-side oracle == CPU oracle
                    ref = oracle_cpu(fr.cpu()).last_hidden_state.mean(dim=1)
                    assert (torch.nn.functional.normalize(ref) - torch.nn.functional.normalize(pooled)).abs().max().item() <= 2e-5
            feats = torch.nn.functional.normalize(pooled)                             # utils.py:44-47
            if shape == "c3":
                lcl_ref = mdf.mif_scores(feats, q[c].cpu())
                idx_ref, st_ref = mdf.mif_select(lcl_ref, K, 1), 0                  # gen_sample.py:87-88
            else:
                idx_ref, st_ref, lcl_ref, _ = mdf.mdf_indices_from_feats(feats, K, W)   # utils.py:55-93
            aux = {"indices": idx_ref, "status": st_ref, "lcl_avg": lcl_ref, "feats": feats}
            eps = (got_lcl[c] - aux["lcl_avg"]).abs().max().item()
            cos = (got_feats[c] * aux["feats"]).sum(dim=1).min().item()
            assert eps <= 1e-3 and cos >= 0.9999, (c, eps, cos)
            max_eps, min_cos = max(max_eps, eps), min(min_cos, cos)
            status_bad += int(int(got_st[c]) != aux["status"])
            first = True
            for a, b in zip(got_idx[c].tolist(), aux["indices"]):
                picks += 1
                if a == b:
                    identical += 1
                    continue
                gap = abs(float(aux["lcl_avg"][a]) - float(aux["lcl_avg"][b]))
                if first:
                    # the FIRST pick that differs must be a tie within the tolerance; on the greedy path it changes the
                    # interval split, so later picks of the clip may differ as its consequence (counted, not excused)
                    assert gap <= 2 * eps, (c, a, b, gap, eps)
                    first = False
                    clips_with_tie += 1
                if gap <= 2 * eps:
                    excused += 1
                else:
                    cascade += 1
    finally:
        enc.close()
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    summary = {"shape": shape, "clips": n_clips, "frames_per_clip": T, "K": K, "W": W, "picks": picks, "identical": identical,
               "excused": excused, "picks_after_a_tie_changed_the_greedy_split": cascade, "clips_with_a_tie": clips_with_tie,
               "status_mismatches": status_bad, "max_eps": max_eps, "min_feature_cosine": min_cos}
    print(f"full-size {shape}:", json.dumps(summary))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, f"parity_{shape}_full_{n_clips}clips.json"), "w") as f:
            json.dump(summary, f, indent=1)
    assert status_bad == 0
    # picks that are not the oracle's, for any reason.  c2 (what the driver runs): measured 0.71 %.  The env-gated shapes tie more
    # often at the same score error (eps ~ 7e-5): c3's relevance scores <feat, q> differ less from frame to frame (measured 4.2 %
    # excused), c4's 512-frame clips hold 32 scenes of near-equal frames and a tie changes the greedy split (measured 1.7 %
    # excused + 3.4 % in its wake, 13 of 64 clips touched) -- all inside the north star's tolerance rule, recorded in profiles/r02
    bound = {"c2": 0.015, "c3": 0.08, "c4": 0.08}[shape]
    assert excused + cascade <= bound * picks, summary
