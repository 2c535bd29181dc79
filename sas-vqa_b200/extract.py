"""The extraction loop of the reference, ``generate_h5_parallel`` (src/preprocessing/extract_features.py:41-111),
over already decoded clips (cv2 decoding is outside the hot path): clips in, the
``sampled_frames`` dataset (+ optional index table) out.

The reference handles one video per iteration -- queue get, sampler call, ``.cpu()``, H5 row write, all serial
(``:80-97``).  Here consecutive clips of one frame size are one batch: equal lengths go through the host-buffer C-ABI
call (H2D, encoder and D2H of consecutive groups overlapped on three streams), different lengths through the ragged
device call (one frame-batched encoder pass, selection kernels on per-clip offsets), and the K frames of every clip
land directly in that clip's row of the dataset.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops, sampler, writer


def generate_h5(clips, model, K: int, W: int, h5_outfile: str, sampling_strategy: str = "repr", group_clips: int = 16,
                debug_counter: dict | None = None, inds_outfile: str | None = None, video_ids=None) -> dict:
    """clips: a sequence of uint8 decoded clips ``[T_i, H, W, 3]`` (RGB, what ``InputGen`` holds before its
    processor call, prefetch_loader.py:57-67).  Writes row i of ``sampled_frames`` for clip i
    (extract_features.py:77-79,96-97) and returns ``dict(debug_counter, indices)``.

    sampling_strategy: 'repr' (MDF, ``sample_representative_frames``), 'uni' (``sample_frames_uniform``) or
    'git6' (``sample_frame_indices(frms, K, 4, len(frms))``) -- extract_features.py:87-94.  Per-clip conditions
    follow the reference: an empty clip stores zero frames ('Zeros'), the top-K fallback counts a 'Failure',
    and T < K on the fallback path raises RuntimeError as ``torch.topk`` does there."""
    if sampling_strategy not in ("repr", "uni", "git6"):
        raise ValueError("sampling_strategy must be one of 'repr', 'uni', 'git6'")
    enc = sampler.as_frame_encoder(model) if sampling_strategy == "repr" else None
    dc = debug_counter if debug_counter is not None else {"Failure": 0, "Zeros": 0}
    n = len(clips)
    all_idx = np.full((n, K), -1, dtype=np.int64)
    with writer.SampledFramesWriter(h5_outfile, n, K) as out:
        if sampling_strategy == "repr":
            # consecutive clips of one frame size are one batch whatever their lengths: equal lengths take the
            # host-buffer call (copies overlapped with compute), different lengths the ragged device call
            i = 0
            while i < n:
                j = i + 1
                while j < n and j - i < group_clips and tuple(clips[j].shape[1:]) == tuple(clips[i].shape[1:]):
                    j += 1
                group = [torch.as_tensor(c) for c in clips[i:j]]
                uniform = all(c.shape[0] == group[0].shape[0] for c in group)
                if uniform and group[0].shape[0] == 0:                              # utils.py:50-52
                    dc["Zeros"] += j - i
                    out[i:j] = np.zeros((j - i, K, 3 * 224 * 224), dtype=np.float32)
                else:
                    if uniform:
                        batch = torch.stack(group)
                        host = batch.pin_memory() if torch.cuda.is_available() else batch
                        res = ops.mdf_sample_host(enc, host, K, W)
                    else:
                        if not torch.cuda.is_available():
                            raise ops._capi.SasvqaError("generate_h5 needs a CUDA device (no CPU fallback)")
                        res = sampler.sample_mdf_ragged([c.cpu() for c in group], enc, K, W)      # host-buffer pipeline
                    st = res["status"]
                    if bool((st == ops.STATUS_TOO_FEW).any()):                       # utils.py:92: topk raises
                        raise RuntimeError("selected index k out of range")
                    dc["Failure"] += int((st == ops.STATUS_FALLBACK).sum())
                    dc["Zeros"] += int((st == ops.STATUS_EMPTY).sum())
                    out[i:j] = res["frames"].reshape(j - i, K, -1)
                    all_idx[i:j] = res["indices"].numpy()
                i = j
        else:
            for i, clip in enumerate(clips):
                clip = torch.as_tensor(clip)
                T = int(clip.shape[0])
                if sampling_strategy == "uni":
                    idx = sampler.uniform_indices(T, K)
                else:
                    idx = sampler.sample_frame_indices(np.arange(T), K, 4, T).tolist()
                sel = clip[torch.as_tensor(idx, dtype=torch.long)]
                if not torch.cuda.is_available():
                    raise ops._capi.SasvqaError("generate_h5 needs a CUDA device (no CPU fallback)")
                dev = sel.cuda()
                if tuple(dev.shape[1:3]) != (224, 224):
                    dev = ops.resize_crop_u8(dev)
                frames = ops.gather_frames_u8(dev.unsqueeze(0), torch.arange(K, dtype=torch.int32, device=dev.device).unsqueeze(0))
                out[i] = frames[0].reshape(K, -1)
                all_idx[i] = idx
    if inds_outfile is not None:
        ids = list(video_ids) if video_ids is not None else [str(i) for i in range(n)]
        writer.write_mdf_inds({v: r for r, v in enumerate(ids)}, all_idx, inds_outfile)
    return dict(debug_counter=dc, indices=all_idx)
