"""Ensemble averaging of member predictions on the device — reference ``average_submissions``
(algonauts2025/grids/average_submissions.py:55-131): per-parcel ``softmax(pearson / temperature)`` weights over the
members (``per_voxel_weights``, :107-110), or one ``softmax(score / temperature)`` weight per member (:111-114), or the
plain mean (:123); applied chunk by chunk to ``(n_timepoints, n_voxels)`` predictions (:117-125).
The one-member-per-GPU variant (all-gather of r + weighted all-reduce) is ``parallel.ensemble_average``."""
from __future__ import annotations

import typing as tp

import numpy as np
import torch

from . import ops
from ._lib import TribeError


def member_weights(pearsons=None, scores=None, per_voxel_weights: bool = True, temperature: float = 1.0, n_voxels: int | None = None,
                   device=None, softmax_over: str = "voxels") -> torch.Tensor:
    """(M, O) fp32 weights on the device.  ``pearsons`` (M, O) per-parcel validation r of every member
    (``pearson.npy``), ``scores`` (M,) their ``val/pearson`` scalars.

    ``softmax_over="voxels"`` reproduces the reference exactly: ``torch.Tensor(pearsons).softmax(dim=1)`` normalises each
    MEMBER's weights over the voxel axis (average_submissions.py:108-109), so a voxel's weights do not sum to one across
    members.  ``"members"`` is the normalisation the variable names suggest (per-voxel convex combination)."""
    if softmax_over not in ("voxels", "members"):
        raise ValueError(softmax_over)
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if per_voxel_weights:
        r = torch.as_tensor(np.asarray(pearsons), dtype=torch.float32).to(dev).contiguous()
        return ops.ensemble_weights(r, temperature, axis=1 if softmax_over == "voxels" else 0)
    s = np.asarray(scores, dtype=np.float64)
    w = np.exp(s / temperature) / np.sum(np.exp(s / temperature))  # average_submissions.py:113
    return torch.as_tensor(w, dtype=torch.float32).to(dev)[:, None].expand(len(s), int(n_voxels)).contiguous()


def average_predictions(preds, weights: torch.Tensor | None) -> torch.Tensor:
    """preds: (M, N, O) stacked member predictions (tensor / array / list of (N, O)); weights (M, O) or None (mean)."""
    if not torch.cuda.is_available():
        raise TribeError("ensemble averaging needs a CUDA device (no CPU fallback)")
    if isinstance(preds, (list, tuple)):
        preds = np.stack([np.asarray(p) for p in preds]) if not torch.is_tensor(preds[0]) else torch.stack(list(preds))
    x = torch.as_tensor(preds, dtype=torch.float32)
    dev = weights.device if weights is not None else torch.device("cuda", torch.cuda.current_device())
    x = x.to(dev).contiguous()
    M = x.shape[0]
    if M <= 48:
        return ops.ensemble_average(x, weights)
    # more members than one launch takes: partial weighted sums (weights already sum to one over ALL members)
    out = None
    for lo in range(0, M, 48):
        hi = min(M, lo + 48)
        w = weights[lo:hi].contiguous() if weights is not None else torch.full((hi - lo, x.shape[2]), 1.0 / M, device=dev)
        part = ops.ensemble_average(x[lo:hi].contiguous(), w)
        out = part if out is None else out.add_(part)
    return out


def average_submissions(predictions: tp.Sequence[dict], pearsons=None, scores=None, weigh_by_score: bool = False,
                        per_voxel_weights: bool = False, temperature: float = 1.0, softmax_over: str = "voxels") -> dict:
    """The arithmetic of ``average_submissions`` on already loaded submissions: ``predictions[m][subject][chunk]`` is the
    (n_timepoints, n_voxels) array of member m.  Returns the same nested dict of averaged float32 numpy arrays."""
    first = predictions[0]
    weights = None
    if weigh_by_score:
        any_chunk = next(iter(next(iter(first.values())).values()))
        weights = member_weights(pearsons, scores, per_voxel_weights, temperature, n_voxels=np.asarray(any_chunk).shape[-1],
                                 softmax_over=softmax_over)
    out: dict = {}
    for sub in first.keys():
        out[sub] = {}
        for chunk in first[sub].keys():
            out[sub][chunk] = average_predictions([data[sub][chunk] for data in predictions], weights).cpu().numpy()
    return out
