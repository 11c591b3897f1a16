"""Window construction and device-side batch assembly — the step BEFORE the hot path.

Reference behaviour mirrored here (host-side integer/float arithmetic, bit-exact index results):
  * ``data_utils/segments.py:144-158``  ``_prepare_strided_windows``: window starts ``arange(start, stop + 1e-8, stride)``;
    ``iter_segments`` (``:161-180``) uses start = timeline start − 4.47 s, stride = duration = 149 s,
    ``drop_incomplete=False``; ``JitterWindows`` (``algonauts2025/callbacks.py:16-44``) adds one random offset per epoch.
  * ``data_utils/base.py:49-53``       ``Frequency.to_ind``: ``int(round(seconds * f))`` (Python banker's rounding).
  * ``data_utils/base.py:167-198``     ``TimedArray._overlap_slice`` and ``__iadd__`` (``:128-165``): a window is a zero
    array of ``max(1, to_ind(duration))`` samples into which the overlapping samples of the timeline's feature array are
    added — i.e. one (dst_start, src_start, length) triple per (window, timeline array).
The copy itself — 13.1 MB per TRIBE window — is ONE gather kernel over device-resident timeline arrays
(``tribe_gather_windows``): no host slicing, no per-window H2D traffic, output directly in the (B, L, D, T) layout
``FmriEncoder`` ingests.
"""
from __future__ import annotations

import typing as tp

import numpy as np
import torch

from . import ops
from ._lib import TribeError
from .segment import SegmentData

LEAD_IN = 4.47  # seconds (3 TRs): data_utils/segments.py:170-171
WINDOW = 149.0  # seconds: stride = duration, data_utils/segments.py:172-173


def to_ind(frequency: float, seconds: float) -> int:
    return int(round(seconds * frequency))


def strided_windows(start: float, stop: float, stride: float = WINDOW, duration: float = WINDOW, drop_incomplete: bool = True):
    eps = 1e-8
    if drop_incomplete:
        stop -= duration
    starts = np.arange(start, stop + eps, stride)
    return starts, np.full_like(starts, fill_value=duration)


def timeline_windows(timeline_start: float, timeline_stop: float, jitter: float = 0.0):
    """Window starts of one timeline exactly as ``iter_segments`` / ``JitterWindows`` enumerate them."""
    return strided_windows(timeline_start - LEAD_IN + jitter, timeline_stop - LEAD_IN + jitter, WINDOW, WINDOW, drop_incomplete=False)


def overlap_slice(arr_start: float, frequency: float, n_samples: int, start: float, duration: float) -> tp.Optional[tuple[int, int]]:
    """``TimedArray(frequency, arr_start, data[..., :n_samples])._overlap_slice(start, duration)`` -> (start_ind, count)."""
    if duration < 0:
        raise ValueError(f"duration should be >=0, got {duration=}")
    arr_duration = n_samples / frequency
    overlap_start = max(start, arr_start)
    overlap_stop = min(start + duration, arr_start + arr_duration)
    if overlap_stop < overlap_start:
        return None
    if overlap_stop == overlap_start and arr_duration and duration:
        return None
    start_ind = to_ind(frequency, overlap_start - arr_start)
    duration_ind = to_ind(frequency, overlap_stop - overlap_start)
    if duration_ind <= 0:
        duration_ind = 1
    if start_ind > n_samples - duration_ind:
        start_ind = n_samples - duration_ind
    if start_ind < 0:
        raise RuntimeError(f"Fail for {start=} {duration=} on array start={arr_start} n={n_samples} f={frequency}")
    return start_ind, duration_ind


def window_triple(win_start: float, win_duration: float, frequency: float, arr_start: float, n_samples: int) -> tuple[int, int, int, int]:
    """(dst_start, src_start, length, window_samples) of the reference's two-stage assembly
    (``data_utils/features/audio.py:236-252`` and the identical text / video / neuro extractors, then ``:104-111``):
      1. ``sub = TimedArray(data, arr_start, f).overlap(win_start, win_duration)`` — a slice of the timeline array whose
         start is re-expressed on the sample grid (``base.py:167-198``);
      2. ``out = TimedArray(f, win_start, win_duration); out += sub`` — both sides sliced again (``base.py:128-165``)."""
    t_win = max(1, to_ind(frequency, win_duration))
    first = overlap_slice(arr_start, frequency, n_samples, win_start, win_duration)
    if first is None:  # the extractor falls back to a 1-sample array at the timeline start, which the window then misses
        return 0, 0, 0, t_win
    s1, n1 = first
    sub_start = s1 / frequency + arr_start
    dst = overlap_slice(win_start, frequency, t_win, sub_start, n1 / frequency)
    src = overlap_slice(sub_start, frequency, n1, win_start, t_win / frequency)
    if dst is None or src is None:
        return 0, 0, 0, t_win
    if dst[1] != src[1]:
        raise ValueError(f"operands could not be broadcast together: window slice {dst} vs array slice {src}")
    return dst[0], s1 + src[0], dst[1], t_win


class TimelineStore:
    """Device-resident per-timeline arrays: ``add(modality, timeline, array (..., T_total), start, frequency)``."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise TribeError("TimelineStore needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.arrays: dict[tuple[str, str], dict] = {}

    def add(self, modality: str, timeline: str, array, start: float, frequency: float) -> None:
        t = torch.as_tensor(array)
        if t.dtype not in (torch.float32, torch.float64):
            t = t.float()
        t = t.to(self.device).contiguous()
        self.arrays[(modality, timeline)] = {"data": t, "start": float(start), "frequency": float(frequency),
                                            "rows_shape": tuple(t.shape[:-1]), "n": int(t.shape[-1])}

    def assemble(self, modality: str, windows: tp.Sequence[tuple[str, float]], duration: float = WINDOW) -> torch.Tensor:
        """windows: (timeline, window start in seconds) per batch element -> fp32 (B, *rows_shape, T) on the device."""
        if not windows:
            raise ValueError("no windows")
        first = self.arrays[(modality, windows[0][0])]
        rows_shape, freq, dtype = first["rows_shape"], first["frequency"], first["data"].dtype
        rows = int(np.prod(rows_shape)) if rows_shape else 1
        triples, ptrs, totals = [], [], []
        t_win = None
        for timeline, w_start in windows:
            a = self.arrays[(modality, timeline)]
            if a["rows_shape"] != rows_shape or a["frequency"] != freq or a["data"].dtype != dtype:
                raise TribeError(f"timeline arrays of modality {modality} disagree in shape / frequency / dtype")
            d0, s0, n, tw = window_triple(float(w_start), duration, freq, a["start"], a["n"])
            t_win = tw if t_win is None else t_win
            triples.append((d0, s0, n))
            ptrs.append(a["data"].data_ptr())
            totals.append(a["n"])
        B = len(windows)
        tri = torch.tensor(triples, dtype=torch.int32).t().contiguous().to(self.device, non_blocking=True)  # (3, B)
        meta = torch.tensor([ptrs, totals], dtype=torch.int64).to(self.device, non_blocking=True)           # (2, B)
        out = torch.empty(B, rows, t_win, device=self.device, dtype=torch.float32)
        ops.gather_windows(meta[0], dtype, meta[1], tri[0], tri[1], tri[2], out)
        return out.view(B, *rows_shape, t_win)

    def batch(self, windows: tp.Sequence[tuple[str, float]], modalities: tp.Sequence[str], subject_ids=None, duration: float = WINDOW) -> SegmentData:
        data = {m: self.assemble(m, windows, duration) for m in modalities}
        if subject_ids is not None:
            data["subject_id"] = torch.as_tensor(subject_ids, dtype=torch.int64).view(-1, 1).to(self.device)
        return SegmentData(data=data, segments=[None] * len(windows))
