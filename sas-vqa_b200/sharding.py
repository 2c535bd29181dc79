"""Multi-GPU layer: videos are independent (the reference loop extract_features.py:80-97 keeps
no state between iterations), so the clip list is sharded by rank -- one process per GPU -- and
the only collective is an all-gather of the per-video index lists at the end.  The reference's
own scheme (intra-video DataParallel over 4 GPUs, extract_features.py:48) is not reproduced.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous balanced slice [start, end) of ``n_items`` owned by ``rank`` (sizes differ by <= 1)."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_by_frames(lengths, world: int) -> list:
    """Contiguous partition of a ragged clip list into ``world`` slices balanced by FRAME count (the encoder's cost is
    per frame, not per clip): boundaries are placed where the running frame total crosses k * total / world, each
    boundary at the clip edge nearest to its target.  Returns ``[(start, end)] * world`` (slices may be empty)."""
    lens = [int(t) for t in lengths]
    n = len(lens)
    total = sum(lens)
    cum = [0]
    for t in lens:
        cum.append(cum[-1] + t)
    bounds = [0]
    for k in range(1, world):
        target = total * k / world
        j = bounds[-1]
        while j < n and cum[j + 1] <= target:
            j += 1
        if j < n and abs(cum[j + 1] - target) < abs(cum[j] - target):
            j += 1                                      # the clip edge nearest to the target
        bounds.append(max(j, bounds[-1]))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def all_gather_ragged_rows(local: torch.Tensor, spans, group=None) -> torch.Tensor:
    """all_gather of per-rank row blocks of DIFFERENT sizes: ``spans[r] = (start, end)`` rows owned by rank r."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    max_rows = max(e - s for s, e in spans)
    pad = torch.zeros((max(max_rows, 1),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([part[: e - s] for part, (s, e) in zip(parts, spans)], dim=0)


def all_gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Concatenate per-rank row blocks (sharded with ``shard_range``) into the full [n_total, ...]
    tensor on every rank.  Works with NCCL (CUDA tensors) and gloo (CPU tensors)."""
    if not dist.is_available() or not dist.is_initialized():
        assert local.shape[0] == n_total
        return local
    world = dist.get_world_size(group)
    if world == 1:
        return local
    max_rows = (n_total + world - 1) // world
    pad = torch.zeros((max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    rows = []
    for r, part in enumerate(parts):
        s, e = shard_range(n_total, r, world)
        rows.append(part[: e - s])
    return torch.cat(rows, dim=0)


def sample_mdf_sharded(make_local_clips, n_clips: int, model, K: int, W: int, group=None, sampler=None) -> dict:
    """Each rank samples its shard of the clip list and all ranks end with the full index table.

    make_local_clips(start, end) -> clips tensor for global clip ids [start, end) (device-resident
    uint8 [n, T, 224, 224, 3] or host tensor).  ``sampler`` defaults to sample_mdf_batch /
    sample_mdf_host by residency.  Returns dict(indices [n_clips, K], status [n_clips], local=...)."""
    from . import sampler as S
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    start, end = shard_range(n_clips, rank, world)
    clips = make_local_clips(start, end)
    if sampler is None:
        sampler = S.sample_mdf_batch if clips.is_cuda else S.sample_mdf_host
    local = sampler(clips, model, K, W)
    idx, status = local["indices"], local["status"]
    if dist.is_initialized() and world > 1 and dist.get_backend(group) == "nccl" and not idx.is_cuda:
        idx, status = idx.cuda(), status.cuda()
    return dict(indices=all_gather_rows(idx, n_clips, group), status=all_gather_rows(status, n_clips, group),
                local=local, shard=(start, end))


def sample_mdf_ragged_sharded(clips, model, K: int, W: int, group=None, sampler=None) -> dict:
    """Ragged clip list (uint8 tensors [T_i, H, W, 3] of one frame size, or a callable ``clips(i)`` plus ``lengths``
    given as ``(lengths, loader)``) sharded by rank in contiguous slices balanced by frame count; every rank ends with
    the full index / status tables.  ``sampler`` defaults to ``sample_mdf_ragged``."""
    from . import sampler as S
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if isinstance(clips, tuple):
        lengths, loader = clips
    else:
        lengths, loader = [int(c.shape[0]) for c in clips], (lambda i: clips[i])
    spans = shard_by_frames(lengths, world)
    start, end = spans[rank]
    run = sampler if sampler is not None else S.sample_mdf_ragged
    if end > start:
        local = run([loader(i) for i in range(start, end)], model, K, W)
        idx, status = local["indices"], local["status"]
    else:
        local = None
        idx, status = torch.zeros(0, K, dtype=torch.int32), torch.zeros(0, dtype=torch.int32)
    if dist.is_initialized() and world > 1 and dist.get_backend(group) == "nccl":
        idx, status = idx.cuda(), status.cuda()
    return dict(indices=all_gather_ragged_rows(idx, spans, group), status=all_gather_ragged_rows(status, spans, group),
                local=local, shard=(start, end), spans=spans)


def generate_inds_sharded(tokenizer, model, qa_samples: list, all_captions: dict, K: int, ds_rate: int = 1,
                          dataset: str = "msvd_qa", group=None, samples_per_call: int = 64) -> list:
    """The MIF step (``gen_sample.py:50-92``) over one split with the QA list sharded by rank: QA samples are
    independent, so each rank scores a contiguous slice with its own scorer replica and the ``[n, K]`` index table is
    all-gathered -- every rank returns the complete ``qa_winds`` list (rank 0 writes it).  No data-path collective."""
    from . import scorer as SC
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n = len(qa_samples)
    start, end = shard_range(n, rank, world)
    local = SC.generate_inds(tokenizer, model, qa_samples[start:end], all_captions, K, ds_rate, dataset=dataset,
                             samples_per_call=samples_per_call)
    idx = torch.tensor([r["sampled_inds"] for r in local], dtype=torch.int32).view(end - start, K)
    if dist.is_initialized() and world > 1 and dist.get_backend(group) == "nccl":
        idx = idx.cuda()
    table = all_gather_rows(idx, n, group).cpu()
    out = []
    for i, sample in enumerate(qa_samples):
        rec = dict(sample)
        rec["sampled_inds"] = [int(v) for v in table[i]]
        out.append(rec)
    return out
