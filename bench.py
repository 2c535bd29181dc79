"""bench.py -- sampled videos/sec of the MDF frame-sampling hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the whole hot path over one batch of synthetic clips:
uint8 decoded frames -> preprocess/patchify -> ViT-B/16 encoder -> pool + L2 norm -> windowed
cosine scores -> greedy selection / top-K fallback -> gather of the K sampled frames.
Workload at every N: BASELINE.json configs[1] per GPU (256 clips x 128 frames, K=16, W=8, bf16
encoder); ranks own disjoint clip ids (weak scaling) and all-gather the index table each step.

`value`  : device-resident inputs (clips already in HBM), CUDA-event timed, max over ranks.
`e2e`    : the same step through the host-buffer C-ABI call (sasvqa_mdf_sample_host): pinned host
           clips in, indices + sampled frames back in host memory, copies inside the timed region.
`roofline`: the dominant kernel (tcgen05 encoder GEMM), timed live with CUDA events on its stream.
`cpu_baseline` / `--impl reference`: the reference's CPU path (HF GitVisionModel fp32 + the oracle
           restatement of its sampler, proven identical in tests/) on the box's host cores, on a
           bounded sample (one clip of the same shape per step).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "sampled videos/sec"
UNIT = "videos/s"
GEMM_FLOP_PER_FRAME = 2 * 196 * 768 * 768 + 12 * (2 * 197 * 768 * (2304 + 768 + 3072 + 3072))   # 33.70 GFLOP
TOTAL_FLOP_PER_FRAME = GEMM_FLOP_PER_FRAME + 12 * (2 * 2 * 197 * 197 * 768)                       # 35.13 GFLOP
# dram__bytes_read.sum + dram__bytes_write.sum of gemm_tcgen05_kernel from one `ncu --set full` capture at
# chunk_frames=2048 (profiles/r01/ncu_gemm_v2_full.txt): qkv 2.43 GB, out_proj 3.04 GB, fc1 3.05 GB, fc2 5.16 GB
# per launch -> mean over the four per-layer launches (the two patch-embed launches per step are negligible).
# algorithmic bytes of the HBM-bound stages, per frame unless noted
HBM_STAGE_BYTES_PER_FRAME = {
    "preprocess": 150528 + 301056,                       # uint8 frame in, bf16 patch matrix out
    "pre_layernorm": 2 * 197 * 768 * 4,                  # fp32 stream read + written in place
    "layernorm": 24 * 197 * (768 * 4 + 768 * 2),         # 24 LayerNorms: fp32 row in, bf16 row out
    "pool_norm": 197 * 768 * 4 + 768 * 4,                # fp32 hidden state in, one unit feature row out
    "attention": 12 * 197 * (2304 * 2 + 768 * 2),        # q|k|v read once, heads written once (12 layers)
}
NCU_GEMM_DRAM_BYTES_PER_LAUNCH = {2048: (2.428e9 + 3.041e9 + 3.047e9 + 5.158e9) / 4}


# Per-GPU shapes of the BASELINE.json configs.  c2 is the metric's configuration (the default, the only one the
# driver runs); the others are reported on request with the same JSON contract.
WORKLOADS = {
    "c1": dict(kind="mdf", clips=1, frames=64, K=16, W=8, H=224, Wd=224,
               desc="MDF sampling of {clips} synthetic {frames}-frame 224x224 clip, K={K}, W={W}: the reference's CPU-runnable case "
                    "(BASELINE configs[0]) as a single-clip latency probe"),
    "c2": dict(kind="mdf", clips=256, frames=128, K=16, W=8, H=224, Wd=224,
               desc="MDF batch of {clips} synthetic {frames}-frame 224x224 clips per GPU, K={K}, W={W} (BASELINE configs[1])"),
    "c2r": dict(kind="mdf-ragged", clips=256, frames=128, K=16, W=8, H=224, Wd=224,
                desc="MDF batch of {clips} synthetic 224x224 clips per GPU of DIFFERENT lengths (uniform in [T/2, 3T/2], T={frames}, "
                     "same total frames as configs[1]) through the ragged call, K={K}, W={W}"),
    "c3": dict(kind="mif", clips=256, frames=128, K=8, W=0, H=224, Wd=224,
               desc="MIF question-conditioned sampling, {clips} clips x {frames} frames per GPU with synthetic question "
                    "embeddings, K={K} (BASELINE configs[2])"),
    "c3x": dict(kind="mif-captions", clips=256, frames=128, K=8, W=0, H=224, Wd=224,
                desc="MIF with the reference's own relevance model: {clips} QA samples x {frames} captions per GPU scored by the "
                     "BERT caption cross-encoder (gen_sample.py:79-88), strided top-K, K={K} (BASELINE configs[2], row f4)"),
    "c4": dict(kind="mdf", clips=64, frames=512, K=32, W=8, H=224, Wd=224,
               desc="MDF long-video sweep, {clips} clips x T={frames} frames per GPU, K={K}, W={W} (BASELINE configs[3])"),
    "c5": dict(kind="mdf+vqa", clips=1250, frames=64, K=16, W=4, H=240, Wd=320,
               desc="end to end: {clips} MSVD/MSRVTT-shaped 240x320 clips x {frames} frames per GPU (10k clips on 8 GPUs) -> "
                    "K0 resize -> MDF K={K}, W={W} -> visual tokens (encoder + GIT visual_projection) of the sampled "
                    "frames (BASELINE configs[4])"),
    "c5x": dict(kind="mdf+vqa-full", clips=512, frames=64, K=16, W=4, H=240, Wd=320,
                desc="end to end with the text side: {clips} MSVD/MSRVTT-shaped 240x320 clips x {frames} frames per GPU -> K0 resize "
                     "-> MDF K={K}, W={W} -> GIT video-QA forward (encoder + visual_projection + 6 decoder blocks + vocabulary "
                     "head, 20 question tokens) on the sampled frames (BASELINE configs[4])"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS),
                    help="BASELINE.json config to run; c2 (the config the metric is quoted on) is what the driver measures")
    ap.add_argument("--clips", type=int, default=None, help="clips per GPU per step (default: the workload's)")
    ap.add_argument("--frames", type=int, default=None, help="frames per clip (T)")
    ap.add_argument("--K", type=int, default=None)
    ap.add_argument("--W", type=int, default=None)
    ap.add_argument("--ds-rate", type=int, default=1, help="MIF stride (workload c3)")
    ap.add_argument("--chunk-frames", type=int, default=2048)
    ap.add_argument("--max-tokens", type=int, default=131072, help="packed tokens per scorer pass (workload c3x)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-clips", type=int, default=1, help="clips in the CPU baseline sample")
    a = ap.parse_args()
    wl = WORKLOADS[a.workload]
    for k in ("clips", "frames", "K", "W"):
        if getattr(a, k) is None:
            setattr(a, k, wl[k])
    a.height, a.width, a.kind, a.desc = wl["H"], wl["Wd"], wl["kind"], wl["desc"]
    return a


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU every 100 ms while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:  # noqa: BLE001
            return local_rank
    return local_rank


# ------------------------------------------------------------------------------------------------
def cpu_reference_setup():
    """The reference's CPU path: HF GitVisionModel (fp32, the third-party encoder it loads) driven
    by the restated sampler (oracle/mdf.py == src/preprocessing/datautils/utils.py:31-94, proven by
    tests/test_oracle_golden.py).  /root/reference itself does not exist on the GPU box."""
    import torch
    from oracle import mdf, vit
    from sasvqa_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth.random_encoder_state_dict(synth.REF_SEED)
    try:
        model = vit.hf_model_from_state_dict(sd)
        enc_name = "HF GitVisionModel fp32"
    except Exception:  # noqa: BLE001  (transformers missing)
        model = vit.VitOracle(sd)
        enc_name = "oracle ViT restatement fp32"
    return mdf, vit, synth, model, enc_name


def cpu_reference_time(args, n_clips: int, repeats: int, warm: int, u8_clips=None):
    """Per-step seconds for the reference CPU sampler on `n_clips` clips/step.  `u8_clips`: the
    very clips the GPU arm sampled (uint8 [n, T, 224, 224, 3] on the host); else generated here."""
    import torch
    mdf, vit, synth, model, enc_name = cpu_reference_setup()
    if u8_clips is None:
        u8_clips = [synth.make_clip(cid, args.frames) for cid in range(n_clips)]
    clips = [vit.image_processor_224(u8_clips[i]) for i in range(n_clips)]
    times, picks = [], None
    with torch.no_grad():
        for it in range(warm + repeats):
            t0 = time.perf_counter()
            for fr in clips:
                _, aux = mdf.sample_representative_frames(fr, model, args.K, args.W, {"Failure": 0, "Zeros": 0},
                                                          return_aux=True)
            dt = time.perf_counter() - t0
            if it >= warm:
                times.append(dt)
            picks = aux["indices"]
    return times, enc_name, torch.get_num_threads(), picks, aux


def run_reference_captions(args):
    """Reference arm of workload c3x: the loop body of generate_inds (gen_sample.py:68-91) one QA sample at a time on the
    host cores -- tokenizer call, HF BertForSequenceClassification forward (fp32; the restated oracle if transformers is
    missing), the reference's strided top-K -- on a bounded sample of the GPU arm's workload."""
    import torch
    from oracle import bert
    from sasvqa_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth.random_scorer_state_dict()
    try:
        from transformers import BertConfig, BertForSequenceClassification
        hf = BertForSequenceClassification(BertConfig(vocab_size=synth.BERT_VOCAB, num_labels=synth.BERT_LABELS)).eval()
        hf.load_state_dict(sd, strict=False)
        model = lambda ids, tts, msk: hf(input_ids=ids, token_type_ids=tts, attention_mask=msk)[0]   # noqa: E731
        name = "HF BertForSequenceClassification fp32"
    except Exception:  # noqa: BLE001
        model = bert.BertScorerOracle(sd)
        name = "oracle BERT restatement fp32"
    tok = synth.SynthTokenizer()
    n = args.cpu_clips
    qa, caps = synth.make_qa_workload(n, args.frames, seed=synth.REF_SEED)
    times = []
    with torch.no_grad():
        for it in range(max(0, min(args.warmup, 1)) + max(1, args.steps)):
            t0 = time.perf_counter()
            bert.generate_inds(tok, model, qa, caps, args.K, args.ds_rate)
            if it >= max(0, min(args.warmup, 1)):
                times.append(time.perf_counter() - t0)
    per_step = sum(times) / len(times)
    value = n / per_step
    sample = (f"{n} QA sample(s) x {args.frames} captions per step (bounded sample of the {args.clips}-sample batch), one sample "
              f"per model call as the reference batches it, {name}, torch.no_grad, {torch.get_num_threads()} threads")
    emit({
        "impl": "reference", "metric": "MIF QA samples/sec through the caption cross-encoder", "value": value,
        "unit": "QA samples/s", "n_gpus": args.gpus, "steps": len(times), "warmup": max(0, min(args.warmup, 1)),
        "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.desc.format(clips=args.clips, frames=args.frames, K=args.K, W=0), "samples_per_gpu": args.clips,
                   "captions_per_sample": args.frames, "K": args.K},
        "cpu_baseline": {"value": value, "unit": "QA samples/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "QA samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.kind == "mif-captions":
        return run_reference_captions(args)
    times, enc_name, threads, _, _ = cpu_reference_time(args, args.cpu_clips, max(1, args.steps),
                                                        max(0, min(args.warmup, 1)))
    per_step = sum(times) / len(times)
    value = args.cpu_clips / per_step
    sample = (f"{args.cpu_clips} clip(s) x {args.frames} frames per step (bounded sample of the {args.clips}-clip batch), "
              f"{enc_name} + restated sampler, torch.no_grad, {threads} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": max(0, min(args.warmup, 1)), "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "frames_per_s": value * args.frames,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(args, n_gpus):
    return {
        "workload": args.desc.format(clips=args.clips, frames=args.frames, K=args.K, W=args.W) +
                    ", random-init ViT-B/16 encoder",
        "clips_per_gpu": args.clips, "frames_per_clip": args.frames, "K": args.K, "W": args.W,
        "global_clips": args.clips * n_gpus, "parallelism": f"dp{n_gpus} (clips sharded by rank)",
        "l2": f"inputs larger than L2 ({args.clips * args.frames * args.height * args.width * 3 / 1e9:.1f} GB uint8 per GPU "
              f"streamed once per step)",
    }


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import sasvqa_b200 as sas
    from sasvqa_b200 import ops, sharding, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (the product has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()

    B, T, K, W = args.clips, args.frames, args.K, args.W
    enc = sas.FrameEncoder(synth.random_encoder_state_dict(synth.REF_SEED), chunk_frames=args.chunk_frames)
    n_total = B * world
    start, end = sharding.shard_range(n_total, rank, world)
    clips = synth.make_clips(range(start, end), T, device=dev, H=args.height, W=args.width)   # uint8, resident in HBM
    q = synth.question_embeddings(range(start, end), device=dev) if args.kind == "mif" else None
    ragged_lengths = None
    if args.kind == "mdf-ragged":                               # same frames, cut into clips of different lengths
        g = torch.Generator().manual_seed(7 + rank)
        lens = torch.randint(T // 2, 3 * T // 2 + 1, (B,), generator=g)
        lens[-1] += B * T - int(lens.sum())                     # keep the total at B * T frames
        while int(lens.min()) < K:                              # (never in practice; keeps every clip samplable)
            lens[int(lens.argmin())] += K
            lens[int(lens.argmax())] -= K
        ragged_lengths = lens.tolist()
        ragged_frames = clips.view(B * T, args.height, args.width, 3)
    dec, qids = None, None
    if args.kind == "mdf+vqa-full":
        dec = sas.GitDecoder(synth.random_git_decoder_state_dict(), max_rows=131072)
        qids = torch.randint(1000, synth.GIT_VOCAB, (B, 20), generator=torch.Generator().manual_seed(5)).to(dev)
        vqa_ev = []
    if args.kind in ("mdf+vqa", "mdf+vqa-full"):
        psd = synth.random_projection_state_dict()
        enc.set_projection(*[psd[f"visual_projection.{k}"] for k in ("0.weight", "0.bias", "1.weight", "1.bias")])
    torch.cuda.synchronize()

    def step():
        if args.kind == "mdf-ragged":
            res = ops.mdf_sample_ragged(enc, ragged_frames, ragged_lengths, K, W, want_frames=True)
        elif args.kind == "mif":
            res = sas.sample_mif_batch(clips, enc, q, K, args.ds_rate, want_frames=True)
            res["status"] = torch.zeros(B, dtype=torch.int32, device=dev)
        else:
            res = sas.sample_mdf_batch(clips, enc, K, W, want_frames=True)
        if args.kind == "mdf+vqa":                               # the downstream forward's visual side, 256 clips at a time
            for b0 in range(0, B, 256):
                res["tokens_probe"] = sas.encode_sampled_frames(res["frames"][b0:b0 + 256], enc)[:, ::197, :8]
        if args.kind == "mdf+vqa-full":                          # the whole downstream forward: next-token logits of the question
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for b0 in range(0, B, 64):
                logits = sas.vqa_logits(res["frames"][b0:b0 + 64], qids[b0:b0 + 64], enc, dec)
                res["answer_probe"] = logits[:, -1, :].argmax(dim=-1)
            e1.record()
            vqa_ev.append((e0, e1))
        table = sharding.all_gather_rows(res["indices"], n_total) if world > 1 else res["indices"]
        return res, table

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        res, table = step()
    barrier()
    launches0 = ops.launch_count()
    enc.profile_enable(True)
    sampler_thread = ClockSampler(physical_gpu_index(local_rank))
    sampler_thread.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        res, table = step()
    ev1.record()
    barrier()
    clocks = sampler_thread.stop()
    elapsed_ms = ev0.elapsed_time(ev1)
    prof = enc.profile_read()
    enc.profile_enable(False)
    launches = ops.launch_count() - launches0
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = n_total / (ms_per_step / 1e3)
    status = res["status"].cpu()

    # ---- roofline of the dominant kernel (all five GEMM shapes run the same tcgen05 kernel)
    gemm_ms = sum(prof[k][0] for k in prof if k.startswith("gemm_"))
    gemm_launches = sum(prof[k][1] for k in prof if k.startswith("gemm_"))
    frames_per_step = B * T + (B * K if args.kind in ("mdf+vqa", "mdf+vqa-full") else 0)   # c5 encodes the K picks a second time
    frames_timed = frames_per_step * args.steps
    achieved_tf = GEMM_FLOP_PER_FRAME * frames_timed / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    stage_ms = {k: round(v[0] / args.steps, 3) for k, v in prof.items()}
    roofline = {
        "bound": "tensor", "kernel": "gemm_tcgen05_kernel", "achieved": achieved_tf, "peak": peaks["tf_sustained"],
        "unit": "TFLOP/s", "frac": achieved_tf / peaks["tf_sustained"],
        "traffic": NCU_GEMM_DRAM_BYTES_PER_LAUNCH.get(args.chunk_frames),
        "traffic_note": "mean DRAM bytes per GEMM launch from ncu (profiles/r01/ncu_gemm_v6_full.txt); algorithmic "
                        "operand+result bytes average 3.4e9 per launch at this chunk size",
        "peak_source": f"{peaks['src']} sustained bf16 (kernel timed inside a long step)",
        "flop_per_launch": GEMM_FLOP_PER_FRAME * frames_timed / max(gemm_launches, 1),
        "avg_launch_ms": gemm_ms / max(gemm_launches, 1), "launches": gemm_launches,
        "gemm_share_of_step": gemm_ms / max(sum(v[0] for v in prof.values()), 1e-9),
        "whole_path_tflops": TOTAL_FLOP_PER_FRAME * frames_per_step / (ms_per_step / 1e3) / 1e12,
        "whole_path_frac_of_sustained": TOTAL_FLOP_PER_FRAME * frames_per_step / (ms_per_step / 1e3) / 1e12 / peaks["tf_sustained"],
        "stage_ms_per_step": stage_ms,
    }

    # achieved GB/s of the HBM-bound stages, from the same CUDA-event scopes
    hbm = {}
    for k, per_frame in HBM_STAGE_BYTES_PER_FRAME.items():
        if prof.get(k, (0.0, 0))[0] > 0:
            gbs = per_frame * frames_timed / (prof[k][0] / 1e3) / 1e9
            hbm[k] = {"gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peaks["hbm"], 3)}
    per_clip = {"scores": T * 768 * 4 + T * 4, "gather": K * (150528 + 602112)}
    for k, per in per_clip.items():
        if prof.get(k, (0.0, 0))[0] > 0:
            gbs = per * B * args.steps / (prof[k][0] / 1e3) / 1e9
            hbm[k] = {"gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peaks["hbm"], 3)}
    if args.kind == "mdf+vqa-full":
        torch.cuda.synchronize()
        timed = vqa_ev[-args.steps:]
        roofline["vqa_forward_ms_per_step"] = sum(a.elapsed_time(b) for a, b in timed) / len(timed)
        roofline["vqa_forward_note"] = ("sas.vqa_logits over the K picks of every clip: encoder + visual_projection + 6 decoder blocks "
                                        "over K*197 + 20 rows per clip + vocabulary head on the text rows; the GEMM roofline above "
                                        "counts only the encoder GEMMs")
    roofline["hbm_stages"] = hbm
    roofline["hbm_peak_gbs"] = peaks["hbm"]

    # ---- end to end through the host-buffer C-ABI call
    e2e = None
    if not args.no_e2e and args.kind == "mdf":
        host_clips = torch.empty(clips.shape, dtype=torch.uint8, pin_memory=True)
        host_clips.copy_(clips)
        idx_h = torch.empty(B, K, dtype=torch.int32, pin_memory=True)
        st_h = torch.empty(B, dtype=torch.int32, pin_memory=True)
        fr_h = torch.empty(B, K, 3, 224, 224, dtype=torch.float32, pin_memory=True)

        def e2e_step():
            out = sas.sample_mdf_host(host_clips, enc, K, W, idx_out=idx_h, status_out=st_h, frames_out=fr_h)
            if world > 1:
                sharding.all_gather_rows(out["indices"].to(dev), n_total)
            return out

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            out = e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_s = float(dt.item()) / args.e2e_steps
        assert torch.equal(out["indices"], res["indices"].cpu()), "host path and device path disagree"
        e2e = {"value": n_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(host_clips.numel()),
               "d2h_bytes_per_step": int(idx_h.numel() * 4 + st_h.numel() * 4 + fr_h.numel() * 4),
               "ms_per_step": e2e_s * 1e3, "steps": args.e2e_steps,
               "api": "sasvqa_mdf_sample_host (pinned uint8 clips in; indices, status and sampled fp32 frames out)"}
        del host_clips, fr_h

    # ---- CPU baseline (rank 0, N=1 only): the reference CPU sampler on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.kind == "mdf" and args.height == 224:
        nc = args.cpu_clips
        times, enc_name, threads, picks, aux = cpu_reference_time(args, nc, 1, 0, u8_clips=clips[:nc].cpu())
        cpu_v = nc / min(times)
        gpu_picks = res["indices"][nc - 1].cpu().tolist()
        # same clip, both arms: identical indices unless the deciding scores tie within the bf16-induced error
        lcl_ref = aux["lcl_avg"]
        gap = max([abs(float(lcl_ref[a]) - float(lcl_ref[b])) for a, b in zip(gpu_picks, picks) if a != b] or [0.0])
        cpu = {"value": cpu_v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{nc} clip(s) x {T} frames: the first clip(s) of the GPU arm's own batch copied to the host, "
                         f"{enc_name} + restated sampler, torch.no_grad, 1 timed pass",
               "cpu_indices": picks, "gpu_indices": gpu_picks, "indices_identical": picks == gpu_picks,
               "max_score_gap_where_different": gap}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, world),
            "frames_per_s": value * T, "clocks": clocks, "gpu_launches": int(launches),
            "status_counts": {"greedy": int((status == 0).sum()), "fallback": int((status == 1).sum())},
            "roofline": roofline,
        }
        if e2e is not None:
            line["e2e"] = e2e
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    if dec is not None:
        dec.close()
    enc.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def run_mif_captions(args):
    """Workload c3x: the MIF step with the reference's cross-encoder.  One step = G QA samples
    x T captions tokenized once up front (the tokenizer is host Python in the reference too, gen_sample.py:80):
    `value` = samples/s with the padded id matrices resident in HBM; `e2e` = the same through
    sasvqa_mif_select_captions_host with the tokenizer's int64 host arrays in and the index table out."""
    import torch
    import torch.distributed as dist
    import sasvqa_b200 as sas
    from sasvqa_b200 import ops, sharding, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (the product has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    G, T, K = args.clips, args.frames, args.K
    sd = synth.random_scorer_state_dict()
    scorer = sas.CaptionScorer(sd, max_tokens=args.max_tokens)
    tok = synth.SynthTokenizer()
    qa, caps = synth.make_qa_workload(G, T, seed=synth.REF_SEED + rank)
    text, pair = [], []
    for smp in qa:
        text += [smp["question"]] * T
        pair += caps[f"video{smp['video']}"]
    batch = tok(text=text, text_pair=pair)
    ids_h, tts_h, msk_h = batch["input_ids"], batch["token_type_ids"], batch["attention_mask"]
    L = int(ids_h.shape[1])
    n_tokens = int(msk_h.sum())
    ids_d, tts_d = ids_h.to(dev, torch.int32), tts_h.to(dev, torch.int32)
    n_total = G * world

    def step():
        logits = scorer.logits(ids_d, tts_d, msk_h)
        idx = ops.topk_strided(logits[:, 0].contiguous().view(G, T), K, args.ds_rate)
        table = sharding.all_gather_rows(idx, n_total) if world > 1 else idx
        return logits, idx, table

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        logits, idx, table = step()
    barrier()
    launches0 = ops.launch_count()
    scorer.profile(True)
    sampler_thread = ClockSampler(physical_gpu_index(local_rank))
    sampler_thread.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        logits, idx, table = step()
    ev1.record()
    barrier()
    clocks = sampler_thread.stop()
    elapsed_ms = ev0.elapsed_time(ev1)
    prof = scorer.profile_read()
    scorer.profile(False)
    launches = ops.launch_count() - launches0
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = n_total / (ms_per_step / 1e3)

    gemm_flop_per_token = 12 * 2 * 768 * (2304 + 768 + 3072 + 3072)
    gemm_ms = sum(v[0] for k, v in prof.items() if k.startswith("gemm_"))
    gemm_launches = sum(v[1] for k, v in prof.items() if k.startswith("gemm_"))
    achieved_tf = gemm_flop_per_token * n_tokens * args.steps / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    ln_bytes = 25 * n_tokens * (768 * 4 * 2 + 768 * 2)          # 25 LayerNorms: fp32 row in, fp32 + bf16 rows out
    roofline = {
        "bound": "tensor", "kernel": "gemm_tcgen05_kernel", "achieved": achieved_tf, "peak": peaks["tf_sustained"],
        "unit": "TFLOP/s", "frac": achieved_tf / peaks["tf_sustained"], "traffic": None,
        "peak_source": f"{peaks['src']} sustained bf16", "launches": gemm_launches,
        "flop_per_launch": gemm_flop_per_token * n_tokens * args.steps / max(gemm_launches, 1),
        "avg_launch_ms": gemm_ms / max(gemm_launches, 1),
        "gemm_share_of_step": gemm_ms / max(sum(v[0] for v in prof.values()), 1e-9),
        "stage_ms_per_step": {k: round(v[0] / args.steps, 3) for k, v in prof.items()},
        "layernorm_gbs": ln_bytes * args.steps / max((prof["layernorm"][0] + prof["embed"][0]) / 1e3, 1e-9) / 1e9,
        "hbm_peak_gbs": peaks["hbm"],
    }

    e2e = None
    if not args.no_e2e:
        idx_h = torch.empty(G, K, dtype=torch.int32, pin_memory=True)
        pin = [x.pin_memory() for x in (ids_h, tts_h, msk_h)]

        def e2e_step():
            out, _ = scorer.select_captions_host(pin[0], pin[1], pin[2], G, K, args.ds_rate, idx_out=idx_h)
            if world > 1:
                sharding.all_gather_rows(out.to(dev), n_total)
            return out

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            out = e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_s = float(dt.item()) / args.e2e_steps
        assert torch.equal(out, idx.cpu()), "host path and device path disagree"
        e2e = {"value": n_total / e2e_s, "unit": "QA samples/s", "h2d_bytes_per_step": int(3 * ids_h.numel() * 8),
               "d2h_bytes_per_step": int(idx_h.numel() * 4), "ms_per_step": e2e_s * 1e3, "steps": args.e2e_steps,
               "api": "sasvqa_mif_select_captions_host (the tokenizer's int64 arrays in, [G, K] caption indices out)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import bert
        torch.set_num_threads(os.cpu_count() or 1)
        oracle_model = bert.BertScorerOracle(sd)
        one = tok(text=text[:T], text_pair=pair[:T])                       # as the reference batches it: one QA sample
        with torch.no_grad():
            t0 = time.perf_counter()
            ref_logits = oracle_model(one["input_ids"], one["token_type_ids"], one["attention_mask"])
            cpu_s = time.perf_counter() - t0
        ref_idx = bert.mif_indices_from_logits(ref_logits, K, args.ds_rate)
        gpu_idx = idx[0].cpu().tolist()
        eps = (ref_logits[:, 0] - logits[:T, 0].cpu()).abs().max().item()
        gap = max([abs(float(ref_logits[a, 0]) - float(ref_logits[b, 0])) for a, b in zip(gpu_idx, ref_idx) if a != b] or [0.0])
        cpu = {"value": 1.0 / cpu_s, "unit": "QA samples/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"1 QA sample x {T} captions (the GPU arm's first sample), fp32 restatement of HF "
                         "BertForSequenceClassification (oracle/bert.py) + the reference's top-K expression, 1 timed pass",
               "cpu_indices": ref_idx, "gpu_indices": gpu_idx, "indices_identical": ref_idx == gpu_idx,
               "max_abs_logit_error": eps, "max_score_gap_where_different": gap}

    if rank == 0:
        line = {
            "metric": "MIF QA samples/sec through the caption cross-encoder", "value": value, "unit": "QA samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": args.desc.format(clips=G, frames=T, K=K, W=0) + ", random-init bert-base-cased classifier, "
                                   "synthetic tokenizer", "samples_per_gpu": G, "captions_per_sample": T, "K": K,
                       "padded_length": L, "tokens_per_step": n_tokens, "mean_tokens_per_pair": n_tokens / (G * T),
                       "max_tokens_per_pass": scorer.max_tokens, "parallelism": f"dp{world} (QA samples sharded by rank)",
                       "l2": f"{n_tokens * 10752 / 1e9:.2f} GB of activations per pass set, larger than L2"},
            "captions_per_s": value * T, "tokens_per_s": n_tokens * world / (ms_per_step / 1e3), "clocks": clocks,
            "gpu_launches": int(launches), "roofline": roofline,
        }
        if e2e is not None:
            line["e2e"] = e2e
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    scorer.close()
    if world > 1:
        dist.destroy_process_group()


def emit(line: dict) -> None:
    """The driver reads ONE JSON line from stdout: everything else (NCCL banners, library printf) was
    re-routed to stderr at start-up, the line goes to the saved real stdout."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    a = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                 # C-level and Python-level stdout -> stderr from here on
    if a.impl == "reference":
        run_reference(a)
    elif a.kind == "mif-captions":
        run_mif_captions(a)
    else:
        run_ours(a)
