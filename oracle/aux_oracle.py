"""TEST INFRASTRUCTURE — CPU restatement (numpy) of the steps either side of the hot path.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this.  Pinned by ``tests/golden/aux_ops.npz``,
which ``oracle/make_golden_aux.py`` generates by RUNNING the reference's own code (PearsonLoss, Rank/TopkAcc, TimedArray,
``_prepare_strided_windows``, ``average_submissions``).

Reference anchors: losses ``modeling_utils/modeling_utils/losses/losses.py:11-42``; retrieval metric
``modeling_utils/modeling_utils/metrics/metrics.py:66-121,194-218``; window slicing ``data_utils/data_utils/base.py:49-53,
128-198``, ``data_utils/data_utils/features/audio.py:100-111,236-252``, ``data_utils/data_utils/segments.py:144-180``;
ensemble ``algonauts2025/grids/average_submissions.py:107-125``; SWA ``algonauts2025/main.py:365-373``."""
from __future__ import annotations

import numpy as np


# ---------------------------------------------------------------------------------------------------- losses
def pearson_loss_and_grad(x: np.ndarray, y: np.ndarray, reduction: str = "mean"):
    """PearsonLoss(dim=1) on (N, O) float64: value and d loss / d x (losses.py:17-42, derivative by hand)."""
    x, y = x.astype(np.float64), y.astype(np.float64)
    xc, yc = x - x.mean(0, keepdims=True), y - y.mean(0, keepdims=True)
    a = (xc * yc).sum(0)
    sx, sy = np.sqrt((xc**2).sum(0)), np.sqrt((yc**2).sum(0))
    D = sx * sy + 1e-8
    loss_p = 1.0 - a / D
    scale = 1.0 / x.shape[1] if reduction == "mean" else 1.0
    grad = -scale * (yc / D - (a * sy / (D**2 * sx)) * xc)
    return (loss_p.mean() if reduction == "mean" else loss_p.sum()), grad


def point_loss_and_grad(x: np.ndarray, y: np.ndarray, kind: str, param: float = 1.0):
    """torch.nn.SmoothL1Loss(beta) / HuberLoss(delta) / L1Loss, reduction='mean' (float64)."""
    d = x.astype(np.float64) - y.astype(np.float64)
    a = np.abs(d)
    if kind == "smooth_l1":
        small = a < param
        l = np.where(small, 0.5 * d * d / max(param, 1e-300), a - 0.5 * param)
        g = np.where(small, d / max(param, 1e-300), np.sign(d))
    elif kind == "huber":
        small = a <= param
        l = np.where(small, 0.5 * d * d, param * (a - 0.5 * param))
        g = np.where(small, d, param * np.sign(d))
    elif kind == "l1":
        l, g = a, np.sign(d)
    else:
        raise ValueError(kind)
    return l.mean(), g / d.size


# ---------------------------------------------------------------------------------------------------- retrieval
def retrieval_scores(x: np.ndarray, y: np.ndarray, eps: float = 1e-15) -> np.ndarray:
    """Rank._compute_sim(norm_kind='y') (metrics.py:83-100)."""
    inv = 1.0 / (eps + np.linalg.norm(y, axis=1))
    return np.einsum("bc,oc,o->bo", x, y, inv)


def retrieval_ranks(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Rank._compute_ranks without labels (metrics.py:102-121)."""
    scores = retrieval_scores(x, y)
    true = np.diag(scores)[:, None]
    with np.errstate(invalid="ignore"):
        gt = np.nansum(scores > true, axis=1)
        ge = np.nansum(scores >= true, axis=1) - 1
    ranks = (gt + ge) / 2.0
    ranks[ranks < 0] = len(scores) // 2
    return ranks


def topk_acc(ranks: np.ndarray, k: int) -> float:
    return float((ranks < k).astype(np.float32).mean())


# ---------------------------------------------------------------------------------------------------- windows
def to_ind(f: float, seconds: float) -> int:
    return int(round(seconds * f))  # base.py:49-53 (banker's rounding)


def _overlap_slice(arr_start, f, n, start, duration):
    """TimedArray._overlap_slice (base.py:167-198) -> (new_start_seconds, start_ind, count) or None."""
    arr_duration = n / f
    o_start, o_stop = max(start, arr_start), min(start + duration, arr_start + arr_duration)
    if o_stop < o_start:
        return None
    if o_stop == o_start and arr_duration and duration:
        return None
    s_ind, d_ind = to_ind(f, o_start - arr_start), to_ind(f, o_stop - o_start)
    if d_ind <= 0:
        d_ind = 1
    if s_ind > n - d_ind:
        s_ind = n - d_ind
    if s_ind < 0:
        raise RuntimeError("negative start index")
    return s_ind / f + arr_start, s_ind, d_ind


def assemble_window(arr: np.ndarray, arr_start: float, f: float, win_start: float, win_duration: float) -> np.ndarray:
    """Extractor flow (audio.py:236-252 then :104-111): sub = array.overlap(window); out = zeros(window); out += sub."""
    t_win = max(1, to_ind(f, win_duration))
    out = np.zeros(arr.shape[:-1] + (t_win,), dtype=arr.dtype)
    first = _overlap_slice(arr_start, f, arr.shape[-1], win_start, win_duration)
    if first is None:
        return out
    sub_start, s1, n1 = first
    sub = arr[..., s1: s1 + n1]
    dst = _overlap_slice(win_start, f, t_win, sub_start, n1 / f)
    src = _overlap_slice(sub_start, f, n1, win_start, t_win / f)
    if dst is None or src is None:
        return out
    out[..., dst[1]: dst[1] + dst[2]] += sub[..., src[1]: src[1] + src[2]]
    return out


def strided_window_starts(start: float, stop: float, stride: float = 149.0, duration: float = 149.0, drop_incomplete: bool = False) -> np.ndarray:
    if drop_incomplete:
        stop -= duration
    return np.arange(start, stop + 1e-8, stride)  # segments.py:144-158


# ---------------------------------------------------------------------------------------------------- ensemble / SWA
def average_members(preds: np.ndarray, pearsons=None, scores=None, weigh_by_score=False, per_voxel_weights=False, temperature=1.0,
                    softmax_axis: int = 1) -> np.ndarray:
    """average_submissions.py:107-125 on one stacked (M, N, O) chunk.  NB the reference's ``pearsons.softmax(dim=1)`` runs
    over the VOXEL axis of the (n_submissions, n_voxels) matrix (softmax_axis=1); axis 0 would be "over members"."""
    preds = np.asarray(preds, dtype=np.float64)
    if not weigh_by_score:
        return preds.mean(0)
    if per_voxel_weights:
        z = np.asarray(pearsons, dtype=np.float64) / temperature
        w = np.exp(z - z.max(softmax_axis, keepdims=True))
        w = (w / w.sum(softmax_axis, keepdims=True))[:, None, :]
    else:
        s = np.asarray(scores, dtype=np.float64)
        w = (np.exp(s / temperature) / np.sum(np.exp(s / temperature)))[:, None, None]
    return (preds * w).sum(0)


def swa_update(avg: np.ndarray, p: np.ndarray, n_averaged: int) -> np.ndarray:
    return avg + (p - avg) / (n_averaged + 1)  # torch.optim.swa_utils.AveragedModel default avg_fn


# ---------------------------------------------------------------------------------------------------- submission assembly
def assemble_submission(batches, labels, target_sample_number, overlap_trs: int = 0):
    """Benchmark.on_test_batch_end / on_test_epoch_end (callbacks.py:56-92) with the evident intent ``overlap_trs = 0``
    (the reference's float 0.0 makes the slice raise for a second window of a chunk).  batches: list of (B, O, T)
    arrays; labels: list of [(subject, chunk)] per batch, already in their final form."""
    sub: dict = {}
    for y, segs in zip(batches, labels):
        for i, (subject, chunk) in enumerate(segs):
            pred = np.asarray(y[i]).T
            per = sub.setdefault(subject, {})
            if chunk not in per:
                per[chunk] = []
            else:
                pred = pred[overlap_trs:]
            per[chunk].append(pred)
    out: dict = {}
    for subject, per in sub.items():
        out[subject] = {}
        for chunk, n in target_sample_number[subject].items():
            result = np.concatenate(per[chunk], axis=0)
            if len(result) < n:
                raise ValueError(f"Warning: {len(result)} predictions for {chunk} but expected at least {n}")
            out[subject][chunk] = result[:n]
    return out
