"""CPU restatement of the SAS-VQA samplers.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Stages follow the reference one to one but return every intermediate (the reference
returns only the gathered frames):

    encode -> pool_norm -> gram -> local_average -> greedy_select | fallback_topk -> gather

Tie rule.  Where the reference's outcome depends on how ``torch.topk`` orders *exactly
equal* scores (CPU ``topk`` is not first-index: ``zeros(8).topk(1)`` gives index 6) the
oracle -- and the CUDA kernels -- take the LOWEST index.  ``argmax`` is first-max in both.
Parity tests excuse a mismatch only when the two picked scores are exactly equal.
"""
from __future__ import annotations

import numpy as np
import torch

CHUNK_FRAMES = 256      # utils.py:29
ADAPTIVE_DIVISOR = 20   # utils.py:30

STATUS_OK = 0
STATUS_FALLBACK = 1     # utils.py:91-93  ('Failure')
STATUS_EMPTY = 2        # utils.py:50-52  ('Zeros')


def resolve_window(W: int, T: int) -> int:
    """utils.py:32-33 -- W == -1 selects the adaptive width T // 20."""
    return T // ADAPTIVE_DIVISOR if W == -1 else W


def pool_norm(model_out) -> torch.Tensor:
    """utils.py:41-47 -- pooler_output if the model has one, else token mean (cls included);
    detach; L2-normalise rows (F.normalize: p=2, dim=1, eps=1e-12)."""
    if hasattr(model_out, "pooler_output"):
        pooled = model_out.pooler_output
    else:
        pooled = model_out.last_hidden_state.mean(dim=1)
    return torch.nn.functional.normalize(pooled.detach())


def encode(frames: torch.Tensor, model) -> torch.Tensor | None:
    """utils.py:35-54 -- serial chunks of 256 frames through the frozen encoder.
    Returns unit-norm features (T, D), or None for an empty clip."""
    T = frames.size(0)
    parts = []
    for s in range(0, T, CHUNK_FRAMES):
        parts.append(pool_norm(model(frames[s:s + CHUNK_FRAMES])))
    if not parts:
        return None
    return torch.cat(parts, dim=0)


def gram(feats: torch.Tensor) -> torch.Tensor:
    """utils.py:55 -- dense T x T cosine-similarity matrix."""
    return feats @ feats.transpose(0, 1)


def local_average(sims: torch.Tensor, W: int) -> torch.Tensor:
    """utils.py:57-61 -- for i in [W, T-W): (sum(S[i, i-W:i+W]) - 1) / (2W - 1); the window
    is asymmetric (includes i-W, excludes i+W) and the self-similarity is assumed to be 1.
    Borders stay exactly 0.  The result lives on the CPU in fp32, like the reference's."""
    T = sims.shape[0]
    out = torch.zeros(T)
    for i in range(W, T - W):
        win = sims[i][i - W:i + W]
        out[i] = (win.sum() - 1) / (len(win) - 1)
    return out


def _first_max(vals: np.ndarray, lo: int, hi: int) -> int:
    """Index of the maximum of vals[lo:hi]; lowest index among exact ties."""
    return lo + int(np.argmax(vals[lo:hi]))


def greedy_select(lcl: torch.Tensor, K: int, W: int) -> list[int]:
    """utils.py:63-88 -- best-first interval splitting.

    Keep the global argmax; then repeatedly take the open interval whose best score is
    highest (ties: smaller left edge, which is what the reference's heap of
    ``(-v, (l, r), idx)`` tuples does), pick its argmax p and split the interval into
    [l, p-W) and [p+W, r) when those are non-empty.  Picks come out in importance order.
    May return fewer than K picks."""
    v = lcl.detach().cpu().numpy().astype(np.float32)
    T = v.shape[0]
    top = _first_max(v, 0, T)
    picks = [top]
    open_ivs: list[tuple[float, int, int, int]] = []   # (-score, l, r, argmax)

    def push(l: int, r: int) -> None:
        p = _first_max(v, l, r)
        open_ivs.append((-float(v[p]), l, r, p))

    if top - W > 0:
        push(0, top - W)
    if top + W < T:
        push(top + W, T)
    while len(picks) < K and open_ivs:
        j = min(range(len(open_ivs)), key=lambda q: (open_ivs[q][0], open_ivs[q][1]))
        _, l, r, p = open_ivs.pop(j)
        picks.append(p)
        if p - W > l:
            push(l, p - W)
        if p + W < r:
            push(p + W, r)
    return picks


def topk_lowest_index(vals, K: int) -> list[int]:
    """Descending top-K with the lowest-index tie rule (torch.topk value order).  -0.0 == 0.0."""
    v = np.asarray(vals, dtype=np.float32)
    if K > v.shape[0]:
        raise RuntimeError("selected index k out of range")   # torch.topk's message
    order = np.lexsort((np.arange(v.shape[0]), -(v + np.float32(0.0))))
    return [int(i) for i in order[:K]]


def mdf_select(lcl: torch.Tensor, K: int, W: int) -> tuple[list[int], int]:
    """utils.py:63-93 -- greedy selection, else discard and take the plain top-K."""
    picks = greedy_select(lcl, K, W)
    if len(picks) < K:
        return topk_lowest_index(lcl.detach().cpu().numpy(), K), STATUS_FALLBACK
    return picks, STATUS_OK


def mdf_indices_from_feats(feats: torch.Tensor, K: int, W: int):
    """a4-a7 given unit-norm features: returns (indices, status, lcl_avg, gram)."""
    W = resolve_window(W, feats.shape[0])
    sims = gram(feats)
    lcl = local_average(sims, W)
    idx, status = mdf_select(lcl, K, W)
    return idx, status, lcl, sims


def sample_representative_frames(frames: torch.Tensor, model, K: int = 16, W: int = 8,
                                 debug_counter: dict | None = None, return_aux: bool = False):
    """utils.py:31-94 with the reference's signature.  aux = dict(indices, status, lcl_avg, feats)."""
    W = resolve_window(W, len(frames))
    feats = encode(frames, model)
    if feats is None:
        debug_counter["Zeros"] += 1
        out = frames.new_zeros(K, 3, 224, 224)
        aux = dict(indices=[], status=STATUS_EMPTY, lcl_avg=None, feats=None)
        return (out, aux) if return_aux else out
    idx, status, lcl, _ = mdf_indices_from_feats(feats, K, W)
    if status == STATUS_FALLBACK:
        debug_counter["Failure"] += 1
    out = frames[torch.as_tensor(idx, dtype=torch.long, device=frames.device)]
    aux = dict(indices=idx, status=status, lcl_avg=lcl, feats=feats)
    return (out, aux) if return_aux else out


def uniform_indices(T: int, K: int) -> list[int]:
    """utils.py:96-109 -- truncating accumulation, not round(i * T / K)."""
    step = T / K
    cur = int(step // 2)
    out = []
    for _ in range(K):
        out.append(cur)
        cur = int(cur + step)
    return out


def git6_indices(T: int, K: int, frame_sample_rate: int = 4, rng=np.random) -> np.ndarray:
    """extract_features.py:32-39 -- random window of K*rate frames ending at a random frame,
    K linspace points clipped to [start, end-1].  Uses numpy's global RNG like the reference
    (seeded 666 at extract_features.py:140) unless another RandomState is passed."""
    span = int(K * frame_sample_rate)
    end = rng.randint(span, T)
    start = end - span
    assert start >= 0
    pts = np.linspace(start, end, num=K)
    return np.clip(pts, start, end - 1).astype(np.int64)


def mif_select(scores, K: int, ds_rate: int = 1) -> list[int]:
    """gen_sample.py:87-88 -- top-K over every ds_rate-th score, mapped back to frame
    indices, best first (never sorted by index)."""
    v = np.asarray(scores, dtype=np.float32)[::ds_rate]
    return [i * ds_rate for i in topk_lowest_index(v, K)]


def mif_scores(feats: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """Embedding-space relevance of BASELINE config 3 (no reference counterpart -- the
    reference's relevance model is the caption cross-encoder of gen_sample.py:80-83):
    scores[t] = <feats[t], q>, fp32."""
    return feats.float() @ q.float()
