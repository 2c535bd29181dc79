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
`cpu_baseline` / `--impl reference`: the reference's OWN sampler (sample_representative_frames, unmodified, from
           oracle/_ref or /root/reference) driving the HF GitVisionModel it loads (fp32) on the box's host cores, on a
           bounded sample of the same clips; timed both as the reference runs it (autograd on) and under no_grad.
`--scaling strong`: --clips is the GLOBAL clip count, sharded over the ranks (weak: --clips per GPU).
At N > 1 rank 0 re-samples 8 clips of the global list unsharded and compares them with the all-gathered table
(`sharding_check`, outside the timed region).
"""
from __future__ import annotations

import argparse
import hashlib
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
NCU_GEMM_SUMMARY = os.path.join(ROOT, "profiles", "r02", "ncu_gemm.json")      # written by tools/summarize_ncu.py
GEMM_SOURCES = ("gemm_tcgen05.cu",)                                              # what the capture's hash covers: the kernel's source file
E2E_PINNED_BYTES_CAP = 9e9                                                        # pinned host memory per rank for the e2e leg


def gemm_source_hash() -> str:
    h = hashlib.sha256()
    for name in GEMM_SOURCES:
        with open(os.path.join(ROOT, "sas-vqa_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def ncu_gemm_traffic(chunk_frames: int):
    """(mean DRAM bytes per GEMM launch, note) from the committed ncu summary -- refused (None) when the summary was
    captured on other GEMM sources or another chunk size than this run's."""
    try:
        summ = json.load(open(NCU_GEMM_SUMMARY))
    except (OSError, ValueError):
        return None, "no ncu summary at profiles/r02/ncu_gemm.json"
    if summ.get("source_sha256_16") != gemm_source_hash():
        return None, (f"profiles/r02/ncu_gemm.json was captured on other GEMM sources (hash {summ.get('source_sha256_16')} != "
                      f"{gemm_source_hash()}): re-capture with tools/gpu_profile.sh")
    if int(summ.get("chunk_frames", 0)) != int(chunk_frames):
        return None, f"ncu summary is for chunk_frames={summ.get('chunk_frames')}, this run uses {chunk_frames}"
    return float(summ["mean_dram_bytes_per_launch"]), (
        f"mean dram__bytes_read+write per per-layer GEMM launch from `ncu --set full` ({summ.get('capture', '?')}); per mode: "
        + ", ".join(f"{k} {v['dram_bytes'] / 1e9:.2f} GB / tensor pipe {v['tensor_pipe_pct']:.0f} %" for k, v in summ["modes"].items()))


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
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --clips per GPU; strong: --clips is the global clip list, sharded over the ranks")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-clips", type=int, default=None,
                    help="clips per GPU per e2e step (default: the step's own batch, capped to ~9 GB of pinned memory)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-clips", type=int, default=2, help="clips in the CPU baseline sample")
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
    """The reference's CPU path: its OWN sample_representative_frames (src/preprocessing/datautils/utils.py:31-94, executed
    unmodified from oracle/_ref -- or /root/reference where that exists) driving the encoder it loads, HF GitVisionModel
    (fp32).  Falls back to the restated sampler (oracle/mdf.py, proven identical in tests/) only when neither tree is there."""
    import torch
    from oracle import mdf, ref_loader, vit
    from sasvqa_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth.random_encoder_state_dict(synth.REF_SEED)
    try:
        model = vit.hf_model_from_state_dict(sd)
        enc_name = "HF GitVisionModel fp32"
    except Exception:  # noqa: BLE001  (transformers missing)
        model = vit.VitOracle(sd)
        enc_name = "oracle ViT restatement fp32"
    if ref_loader.available():
        fn, _ = ref_loader.load_sampler_fns()
        kind, fn_name = "reference", f"the reference's own sample_representative_frames ({ref_loader.source_kind()})"
    else:
        fn = lambda fr, m, K, W, dc: mdf.sample_representative_frames(fr, m, K, W, dc)   # noqa: E731
        kind, fn_name = "port", "restated sampler (oracle/mdf.py)"
    return dict(fn=fn, kind=kind, fn_name=fn_name, model=model, enc_name=enc_name, mdf=mdf, vit=vit, synth=synth)


def picked_indices(frames, picked):
    """Indices of the returned frames inside the clip (the reference returns frames, never indices)."""
    flat = frames.flatten(1)
    out = []
    for f in picked.flatten(1):
        hit = (flat == f).all(dim=1).nonzero().flatten().tolist()
        out.append(hit[0] if hit else -1)
    return out


def cpu_reference_time(args, n_clips: int, repeats: int, warm: int, u8_clips=None):
    """Times the reference CPU sampler on `n_clips` clips per step, `repeats` timed steps after `warm` untimed ones, in
    BOTH modes of BASELINE.md section 4: exactly as the reference runs it (no torch.no_grad, utils.py:39-48) and under
    torch.no_grad().  `u8_clips`: the very clips the GPU arm sampled (uint8 [n, T, H, W, 3] on the host); else generated
    here.  Returns dict(times_grad, times_no_grad, picks (of the last clip), ...)."""
    import torch
    ctx = cpu_reference_setup()
    synth, vit = ctx["synth"], ctx["vit"]
    if u8_clips is None:
        u8_clips = [synth.make_clip(cid, args.frames, H=args.height, W=args.width) for cid in range(n_clips)]
    if args.height == 224 and args.width == 224:
        clips = [vit.image_processor_224(u8_clips[i]) for i in range(n_clips)]
        proc = "CLIPImageProcessor arithmetic (rescale + normalise; 224x224 input)"
    else:
        try:                                                    # what the reference calls (prefetch_loader.py:74-75)
            from transformers import CLIPImageProcessor
            hf_proc = CLIPImageProcessor()
            clips = [hf_proc(images=list(u8_clips[i].numpy()), return_tensors="pt")["pixel_values"] for i in range(n_clips)]
            proc = "HF CLIPImageProcessor (bicubic shortest-edge resize + centre crop + rescale + normalise), untimed"
        except Exception:  # noqa: BLE001
            from oracle import resize
            clips = [vit.image_processor_224(torch.from_numpy(resize.resize_crop_u8(u8_clips[i].numpy()))) for i in range(n_clips)]
            proc = "image processor restatement (oracle/resize.py), untimed"
    out = {"times_grad": [], "times_no_grad": [], "kind": ctx["kind"], "threads": torch.get_num_threads(),
           "what": f"{ctx['fn_name']} + {ctx['enc_name']}; frames from the {proc}"}
    for mode in ("grad", "no_grad"):
        with (torch.no_grad() if mode == "no_grad" else torch.enable_grad()):
            for it in range(warm + repeats):
                t0 = time.perf_counter()
                for fr in clips:
                    picked = ctx["fn"](fr, ctx["model"], args.K, args.W, {"Failure": 0, "Zeros": 0})
                dt = time.perf_counter() - t0
                if it >= warm:
                    out["times_" + mode].append(dt)
    out["picks"] = picked_indices(clips[-1], picked)
    with torch.no_grad():     # the oracle's scores of the same clip (to size a tie when the two arms differ)
        _, aux = ctx["mdf"].sample_representative_frames(clips[-1], ctx["model"], args.K, args.W, {"Failure": 0, "Zeros": 0},
                                                         return_aux=True)
    out["lcl_avg"] = aux["lcl_avg"]
    return out


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
    """`--impl reference`: the reference's own CPU implementation of the path on this box's host cores, rank 0 only.
    One step = `--cpu-clips` clips of the workload's shape (a bounded sample of the batch); `--steps` timed steps after at
    most one warm-up step, in both modes; `value` = clips / best step of the FASTER mode (the conservative denominator)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.kind == "mif-captions":
        return run_reference_captions(args)
    warm = max(0, min(args.warmup, 1))
    r = cpu_reference_time(args, args.cpu_clips, max(1, args.steps), warm)
    v_grad = args.cpu_clips / min(r["times_grad"])
    v_nograd = args.cpu_clips / min(r["times_no_grad"])
    value = max(v_grad, v_nograd)
    per_step = args.cpu_clips / value
    sample = (f"{args.cpu_clips} clip(s) x {args.frames} frames per step (bounded sample of the {args.clips}-clip batch), "
              f"{r['what']}, best of {len(r['times_grad'])} steps per mode after {warm} warm-up, {r['threads']} threads")
    if args.kind in ("mdf+vqa", "mdf+vqa-full", "mif"):
        sample += "; the CPU arm runs the MDF sampler only (no downstream forward / question scoring), i.e. LESS work than the GPU arm"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(r["times_grad"]),
        "warmup": warm, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "frames_per_s": value * args.frames,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["threads"], "kind": r["kind"], "sample": sample,
                         "as_reference_runs_it": {"value": v_grad, "note": "no torch.no_grad (utils.py:39-48 builds autograd graphs)",
                                                  "step_s": [round(t, 3) for t in r["times_grad"]]},
                         "under_no_grad": {"value": v_nograd, "step_s": [round(t, 3) for t in r["times_no_grad"]]},
                         "value_is": "the faster of the two modes", "host_cpu_count": os.cpu_count(),
                         "indices_last_clip": r["picks"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(args, n_gpus):
    n_total = args.clips if args.scaling == "strong" else args.clips * n_gpus
    per_gpu = (n_total + n_gpus - 1) // n_gpus
    return {
        "workload": args.desc.format(clips=per_gpu, frames=args.frames, K=args.K, W=args.W) +
                    ", random-init ViT-B/16 encoder",
        "clips_per_gpu": per_gpu, "frames_per_clip": args.frames, "K": args.K, "W": args.W,
        "global_clips": n_total, "parallelism": f"dp{n_gpus} (clips sharded by rank)",
        "l2": f"inputs larger than L2 ({per_gpu * args.frames * args.height * args.width * 3 / 1e9:.1f} GB uint8 per GPU "
              f"streamed once per step)",
    }


# ------------------------------------------------------------------------------------------------
class EnergyMeter:
    """NVML total-energy counter of one GPU (mJ since driver load): joules over a region = the difference."""

    def __init__(self, index: int):
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.read()
        except Exception:  # noqa: BLE001
            self.h = None

    def read(self):
        return self.nv.nvmlDeviceGetTotalEnergyConsumption(self.h) / 1e3 if self.h is not None else None


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

    T, K, W = args.frames, args.K, args.W
    vqa = args.kind in ("mdf+vqa", "mdf+vqa-full")
    enc = sas.FrameEncoder(synth.random_encoder_state_dict(synth.REF_SEED), chunk_frames=args.chunk_frames)
    n_total = args.clips if args.scaling == "strong" else args.clips * world
    start, end = sharding.shard_range(n_total, rank, world)
    B = end - start                                                  # this rank's clips (strong: differs by <= 1 over ranks)
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
    dec, qids_all = None, None
    vqa_ev = []
    if args.kind == "mdf+vqa-full":
        dec = sas.GitDecoder(synth.random_git_decoder_state_dict(), max_rows=131072)
        qids_all = torch.randint(1000, synth.GIT_VOCAB, (n_total, 20), generator=torch.Generator().manual_seed(5))
        qids = qids_all[start:end].to(dev)
    if vqa:
        psd = synth.random_projection_state_dict()
        enc.set_projection(*[psd[f"visual_projection.{k}"] for k in ("0.weight", "0.bias", "1.weight", "1.bias")])
    torch.cuda.synchronize()

    def downstream(frames_dev, q_ids, res, record):
        """configs[4]: the video-QA forward on the sampled frames (visual side only for c5, the whole forward for c5x)"""
        n = frames_dev.shape[0]
        if args.kind == "mdf+vqa":
            for b0 in range(0, n, 256):
                res["tokens_probe"] = sas.encode_sampled_frames(frames_dev[b0:b0 + 256], enc)[:, ::197, :8]
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            answers = []
            for b0 in range(0, n, 64):
                logits = sas.vqa_logits(frames_dev[b0:b0 + 64], q_ids[b0:b0 + 64], enc, dec)
                answers.append(logits[:, -1, :].argmax(dim=-1))
            res["answer_probe"] = torch.cat(answers)
            e1.record()
            if record:
                vqa_ev.append((e0, e1))

    def step():
        if args.kind == "mdf-ragged":
            res = ops.mdf_sample_ragged(enc, ragged_frames, ragged_lengths, K, W, want_frames=True)
        elif args.kind == "mif":
            res = sas.sample_mif_batch(clips, enc, q, K, args.ds_rate, want_frames=True)
            res["status"] = torch.zeros(B, dtype=torch.int32, device=dev)
        else:
            res = sas.sample_mdf_batch(clips, enc, K, W, want_frames=True)
        if vqa:
            downstream(res["frames"], qids if dec is not None else None, res, True)
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
    gpu_index = physical_gpu_index(local_rank)
    sampler_thread = ClockSampler(gpu_index)
    energy = EnergyMeter(gpu_index)
    sampler_thread.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    j0 = energy.read()
    ev0.record()
    for _ in range(args.steps):
        res, table = step()
    ev1.record()
    barrier()
    j1 = energy.read()
    clocks = sampler_thread.stop()
    elapsed_ms = ev0.elapsed_time(ev1)
    prof = enc.profile_read()
    enc.profile_enable(False)
    launches = ops.launch_count() - launches0
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        # every rank's own time, SM clock under the power cap and joules: the step is the MAX over ranks, so the slowest
        # piece of silicon on the box sets the number (GPUs differ by 10-15 % in clock at the same 1 kW cap)
        mine = torch.tensor([elapsed_ms / args.steps, float(clocks["sm_mhz"] or 0.0),
                             ((j1 - j0) / args.steps) if j0 is not None and j1 is not None else 0.0], dtype=torch.float64, device=dev)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"ms_per_step": [round(float(x[0]), 2) for x in allr], "sm_mhz": [float(x[1]) for x in allr],
                    "joules_per_step": [round(float(x[2]), 1) for x in allr]}
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = n_total / (ms_per_step / 1e3)
    status = res["status"].cpu()
    idx_cpu = res["indices"].cpu()

    # ---- roofline of the dominant kernel (all five GEMM shapes run the same tcgen05 kernel)
    gemm_ms = sum(prof[k][0] for k in prof if k.startswith("gemm_"))
    gemm_launches = sum(prof[k][1] for k in prof if k.startswith("gemm_"))
    frames_per_step = B * T + (B * K if vqa else 0)                  # c5 encodes the K picks a second time
    frames_timed = frames_per_step * args.steps
    achieved_tf = GEMM_FLOP_PER_FRAME * frames_timed / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    stage_ms = {k: round(v[0] / args.steps, 3) for k, v in prof.items()}
    traffic, traffic_note = ncu_gemm_traffic(enc.chunk_frames)
    roofline = {
        "bound": "tensor", "kernel": "gemm_tcgen05_kernel", "achieved": achieved_tf, "peak": peaks["tf_sustained"],
        "unit": "TFLOP/s", "frac": achieved_tf / peaks["tf_sustained"],
        "traffic": traffic, "traffic_note": traffic_note,
        "peak_source": f"{peaks['src']} sustained bf16 (kernel timed inside a long step)",
        "flop_per_launch": GEMM_FLOP_PER_FRAME * frames_timed / max(gemm_launches, 1),
        "avg_launch_ms": gemm_ms / max(gemm_launches, 1), "launches": gemm_launches,
        "gemm_share_of_step": gemm_ms / max(sum(v[0] for v in prof.values()), 1e-9),
        "whole_path_tflops": TOTAL_FLOP_PER_FRAME * frames_per_step / (ms_per_step / 1e3) / 1e12,
        "whole_path_frac_of_sustained": TOTAL_FLOP_PER_FRAME * frames_per_step / (ms_per_step / 1e3) / 1e12 / peaks["tf_sustained"],
        "stage_ms_per_step": stage_ms,
    }
    if j0 is not None and j1 is not None:
        joules = (j1 - j0) / args.steps
        roofline["energy"] = {"joules_per_step": round(joules, 1), "avg_power_w": round(joules / (ms_per_step / 1e3), 1),
                              "pj_per_flop_whole_path": round(joules / (TOTAL_FLOP_PER_FRAME * frames_per_step) * 1e12, 4),
                              "source": "nvmlDeviceGetTotalEnergyConsumption around the timed region (this rank's GPU)"}

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

    # ---- fallback clips whose picks include exact-0.0 border scores (lcl_avg[i < W or i >= T - W] == 0, utils.py:57-61): there
    # the order among equal scores is torch.topk's on the CPU and lowest-index-first here, so the stored frames may differ
    st_counts = {"greedy": int((status == 0).sum()), "fallback": int((status == 1).sum()), "empty": int((status == 2).sum()),
                 "too_few": int((status == 3).sum())}
    if args.kind in ("mdf", "mdf+vqa", "mdf+vqa-full"):
        Wr = T // 20 if W == -1 else W
        border = (idx_cpu < Wr) | (idx_cpu >= T - Wr)
        st_counts["fallback_zero_border_clips"] = int(((status == 1) & border.any(dim=1)).sum())

    # ---- sharding invariance (N > 1): rank 0 re-samples 8 clips of the GLOBAL list on its own and compares with the gathered table
    sharding_check = None
    if world > 1 and rank == 0:
        if args.kind == "mdf-ragged":
            sharding_check = {"skipped": "ragged clips are cut from each rank's own frame stream; see tests/test_gpu_multi.py"}
        else:
            ids = sorted({int(round(i * (n_total - 1) / 7)) for i in range(8)})
            probe = synth.make_clips(ids, T, device=dev, H=args.height, W=args.width)
            if args.kind == "mif":
                alone = sas.sample_mif_batch(probe, enc, synth.question_embeddings(ids, device=dev), K, args.ds_rate)["indices"]
            else:
                alone = sas.sample_mdf_batch(probe, enc, K, W, want_frames=False)["indices"]
            same = torch.equal(alone.cpu(), table.cpu()[torch.tensor(ids)])
            sharding_check = {"clips": len(ids), "clip_ids": ids, "identical": bool(same),
                              "how": "rank 0 regenerates these global clips, samples them unsharded (one 8-clip call) and compares "
                                     "the indices with the rows of the all-gathered table (outside the timed region)"}
            del probe

    # ---- end to end through the host-buffer C-ABI calls: pinned host inputs, results back in host memory, every copy timed
    e2e = None
    if not args.no_e2e:
        bytes_per_clip = T * args.height * args.width * 3 + K * ops.FRAME_ELEMS * 4
        nb = args.e2e_clips or max(1, min(B, int(E2E_PINNED_BYTES_CAP // bytes_per_clip)))
        nb = min(nb, B)
        if world > 1 and args.scaling == "strong":                   # every rank the same e2e batch so the global count is exact
            t_nb = torch.tensor([nb], device=dev)
            dist.all_reduce(t_nb, op=dist.ReduceOp.MIN)
            nb = int(t_nb.item())
        host_clips = torch.empty((nb,) + tuple(clips.shape[1:]), dtype=torch.uint8, pin_memory=True)
        host_clips.copy_(clips[:nb])
        idx_h = torch.empty(nb, K, dtype=torch.int32, pin_memory=True)
        st_h = torch.empty(nb, dtype=torch.int32, pin_memory=True)
        fr_h = torch.empty(nb, K, 3, 224, 224, dtype=torch.float32, pin_memory=True)
        h2d, d2h = int(host_clips.numel()), int(idx_h.numel() * 4 + st_h.numel() * 4 + fr_h.numel() * 4)
        if args.kind == "mdf-ragged":
            cum, nclip = 0, 0
            while nclip < B and cum + ragged_lengths[nclip] <= nb * T:
                cum += ragged_lengths[nclip]
                nclip += 1
            host_frames = host_clips.view(nb * T, args.height, args.width, 3)[:cum]
            api = "sasvqa_mdf_sample_ragged_host"
            n_e2e, h2d = nclip, int(host_frames.numel())
            d2h = n_e2e * (K * 4 + 4 + K * ops.FRAME_ELEMS * 4)
        elif args.kind == "mif":
            q_h = q[:nb].cpu().pin_memory()
            api, n_e2e = "sasvqa_mif_sample_host_hw", nb
            h2d += int(q_h.numel() * 4)
            d2h -= int(st_h.numel() * 4)
        else:
            api, n_e2e = "sasvqa_mdf_sample_host_hw", nb
        if vqa:
            ans_h = torch.empty(nb, dtype=torch.int64, pin_memory=True)
            qids_h = qids_all[start:start + nb].to(torch.int32).pin_memory() if dec is not None else None
            api += " -> sampled frames back to the GPU in batches -> " + ("sasvqa_git_vqa_logits_f32" if dec is not None else
                                                                            "sasvqa_visual_tokens_f32")
            h2d += int(fr_h.numel() * 4) + (int(qids_h.numel() * 4) if qids_h is not None else 0)
            d2h += int(ans_h.numel() * 8) if dec is not None else 0

        def e2e_step():
            if args.kind == "mdf-ragged":
                out = ops.mdf_sample_ragged_host(enc, host_frames, ragged_lengths[:n_e2e], K, W, idx_out=idx_h[:n_e2e],
                                                 status_out=st_h[:n_e2e], frames_out=fr_h[:n_e2e])
            elif args.kind == "mif":
                out = sas.sample_mif_host(host_clips, enc, q_h, K, args.ds_rate, idx_out=idx_h, frames_out=fr_h, want_frames=True)
            else:
                out = sas.sample_mdf_host(host_clips, enc, K, W, idx_out=idx_h, status_out=st_h, frames_out=fr_h)
            if vqa:        # the consumer reads the stored frames back (the reference couples the two programs through the H5 file)
                r2 = {}
                for b0 in range(0, nb, 64):
                    fr_d = fr_h[b0:b0 + 64].to(dev, non_blocking=True)
                    qd = qids_h[b0:b0 + 64].to(dev, non_blocking=True) if qids_h is not None else None
                    downstream(fr_d, qd, r2, False)
                    if dec is not None:
                        ans_h[b0:b0 + 64].copy_(r2["answer_probe"], non_blocking=True)
                torch.cuda.synchronize()
            if world > 1:
                sharding.all_gather_rows(out["indices"].to(dev), n_e2e * world)
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
        assert torch.equal(out["indices"][:n_e2e], idx_cpu[:n_e2e]), "host path and device path disagree"
        e2e = {"value": n_e2e * world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": e2e_s * 1e3, "steps": args.e2e_steps, "clips_per_gpu_per_step": n_e2e,
               "api": api + " (pinned host buffers in; indices, status and sampled fp32 frames out)",
               "indices_equal_device_path": True}
        del host_clips, fr_h

    # ---- CPU baseline (rank 0, N=1 only): the reference's own sampler on a bounded sample of this very batch
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.kind in ("mdf", "mdf+vqa", "mdf+vqa-full"):
        nc = max(1, min(args.cpu_clips, B))
        r = cpu_reference_time(args, nc, 1, 0, u8_clips=clips[:nc].cpu())
        v_grad, v_nograd = nc / min(r["times_grad"]), nc / min(r["times_no_grad"])
        picks = r["picks"]
        gpu_picks = idx_cpu[nc - 1].tolist()
        # same clip, both arms: identical indices unless the deciding scores tie within the bf16-induced error
        lcl_ref = r["lcl_avg"]
        gap = max([abs(float(lcl_ref[a]) - float(lcl_ref[b])) for a, b in zip(gpu_picks, picks) if a != b and a >= 0 and b >= 0]
                  or [0.0])
        cpu = {"value": max(v_grad, v_nograd), "unit": UNIT, "cores": r["threads"], "kind": r["kind"],
               "sample": f"{nc} clip(s) x {T} frames: the first clip(s) of the GPU arm's own batch copied to the host, "
                         f"{r['what']}, 1 timed pass per mode",
               "as_reference_runs_it": {"value": v_grad, "note": "no torch.no_grad (utils.py:39-48 builds autograd graphs)"},
               "under_no_grad": {"value": v_nograd}, "value_is": "the faster of the two modes",
               "host_cpu_count": os.cpu_count(),
               "cpu_indices": picks, "gpu_indices": gpu_picks, "indices_identical": picks == gpu_picks,
               "max_score_gap_where_different": gap}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, world),
            "frames_per_s": value * T, "clocks": clocks, "gpu_launches": int(launches),
            "status_counts": st_counts, "roofline": roofline,
        }
        if per_rank is not None:
            line["per_rank"] = per_rank
        if sharding_check is not None:
            line["sharding_check"] = sharding_check
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
