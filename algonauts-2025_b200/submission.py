"""Submission assembly — the step AFTER the hot path (reference ``Benchmark`` callback, algonauts2025/callbacks.py:47-103):
test-time predictions (B, 1000, T) are transposed per window to (T, 1000), appended to their (subject, chunk) list in
arrival order, concatenated, checked against and truncated to the expected number of fMRI samples.  Here the windows
stay on the device until ``finalize``; the transposes are one kernel launch per batch (``tribe_transpose_last2``) and
each chunk costs one device concatenation + one D2H copy instead of a ``.cpu().numpy()`` per window."""
from __future__ import annotations

import typing as tp

import numpy as np
import torch

from . import ops
from ._lib import TribeError


class SubmissionAssembler:
    def __init__(self) -> None:
        self.submission: dict[str, dict[str, list[torch.Tensor]]] = {}

    def reset(self) -> None:  # on_test_epoch_start
        self.submission = {}

    @torch.no_grad()
    def add_batch(self, y_pred: torch.Tensor, subjects: tp.Sequence[str], chunks: tp.Sequence[str], overlap_trs: int = 0) -> None:
        """``y_pred`` (B, O, T) on the device; ``subjects[i]`` / ``chunks[i]`` as the reference derives them from the
        window's events (``"…/sub-01"`` -> ``"sub-01"``, ``"…:e01a"`` -> ``"s07e01a"``, callbacks.py:62-68)."""
        if not y_pred.is_cuda:
            raise TribeError("SubmissionAssembler needs CUDA predictions (no CPU fallback)")
        if not (len(subjects) == len(chunks) == y_pred.shape[0]):
            raise ValueError("one subject and one chunk label per window")
        pred_t = ops.transpose_last2(y_pred.detach().float().contiguous())  # (B, T, O): every window's pred.T
        for i, (subject, chunk) in enumerate(zip(subjects, chunks)):
            subject = subject.split("/")[1] if "/" in subject else subject
            chunk = "s07" + chunk.split(":")[1] if ":" in chunk else chunk
            per_subject = self.submission.setdefault(subject, {})
            pred = pred_t[i]
            if chunk not in per_subject:
                per_subject[chunk] = []
            else:
                pred = pred[overlap_trs:]  # remove the overlap except on the first window of a chunk (callbacks.py:73-74)
            per_subject[chunk].append(pred)

    def finalize(self, target_sample_number: dict[str, dict[str, int]]) -> dict[str, dict[str, np.ndarray]]:
        """``target_sample_number[subject][chunk]`` = expected fMRI samples (callbacks.py:80-92).  Returns
        ``{subject: {chunk: float32 (n_samples, O)}}`` exactly like ``Benchmark.submission_dict`` after the epoch."""
        out: dict[str, dict[str, np.ndarray]] = {}
        for subject, chunks in self.submission.items():
            out[subject] = {}
            for chunk, sample_number in target_sample_number[subject].items():
                result = torch.cat(chunks[chunk], dim=0)
                if len(result) < sample_number:
                    raise ValueError(f"Warning: {len(result)} predictions for {chunk} but expected at least {sample_number}")
                out[subject][chunk] = result[:sample_number].cpu().numpy()
        return out
