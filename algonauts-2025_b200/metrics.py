"""Evaluation metrics of the hot path on the Pearson-statistics kernels.

Mirrors (reference paths): ``MultidimPearsonCorrCoef`` and ``GroupedMetric`` modeling_utils/modeling_utils/metrics/
base.py:26-29, 39-91 (torchmetrics ``PearsonCorrCoef`` semantics, oracle/tm_pearson.py); ``compute_multidim_pearson``
algonauts2025/main.py:459-477 (the per-parcel ``scipy.stats.pearsonr`` loop).  State = per-parcel fp64 sufficient
statistics [n, Σx, Σy, Σx², Σy², Σxy] accumulated on the device by ``tribe_pearson_stats``; under ``torch.distributed``
the statistics are all-reduced (one 48 KB message) instead of torchmetrics' state gather."""
from __future__ import annotations

import typing as tp

import numpy as np
import torch
from torch import nn

from . import ops
from ._lib import TribeError


def _world_size() -> int:
    import torch.distributed as dist

    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _merged_stats(stats: torch.Tensor, shift: torch.Tensor | None) -> torch.Tensor:
    """A copy of the local statistics ready for ``pearson_finalize``: under ``torch.distributed`` every rank's block is
    re-expressed about pivot 0 in fp64 (ranks picked their pivots from their own first rows) and summed with one
    all-reduce (48 KB for 1000 parcels) — the reduction torchmetrics does by gathering its states."""
    st = stats.clone()
    if _world_size() > 1:
        import torch.distributed as dist

        if shift is not None:
            ops.pearson_recenter(st, shift, None)
        dist.all_reduce(st, op=dist.ReduceOp.SUM)
    return st


class MultidimPearsonCorrCoef(nn.Module):
    """``torchmetrics.PearsonCorrCoef(num_outputs=O)`` whose ``compute()`` returns the mean over outputs."""

    is_differentiable = False
    higher_is_better = True

    def __init__(self, num_outputs: int = 1, **kwargs: tp.Any) -> None:
        super().__init__()
        self.num_outputs = num_outputs
        self.register_buffer("stats", torch.zeros(1, 6, num_outputs, dtype=torch.float64), persistent=False)
        # per-parcel pivots (first sample seen since the last reset): the kernels accumulate moments of (x - pivot), which
        # stay well conditioned when |mean| >> sigma (torchmetrics / scipy centre before multiplying)
        self.register_buffer("shift", torch.zeros(2, num_outputs, dtype=torch.float32), persistent=False)
        self._has_shift = False

    def _dev(self, like: torch.Tensor):
        if not like.is_cuda:
            raise TribeError("Pearson metric needs CUDA tensors (no CPU fallback)")
        if self.stats.device != like.device:
            self.stats, self.shift = self.stats.to(like.device), self.shift.to(like.device)

    def _update(self, preds, target, layout, **kw):
        self._dev(preds)
        preds, target = preds.detach().float().contiguous(), target.detach().to(preds.device).float().contiguous()
        if preds.numel() == 0:
            return
        if not self._has_shift:
            ops.pearson_pick_shift(preds, target, self.shift, layout=layout)
            self._has_shift = True
        ops.pearson_stats(preds, target, self.stats, layout=layout, shift=self.shift, **kw)

    def update(self, preds: torch.Tensor, target: torch.Tensor) -> None:
        """preds/target: (N, O) — the flattened ``(b t) d`` matrices of pl_module.py:54-55."""
        if preds.ndim == 1:
            preds, target = preds[:, None], target[:, None]
        self._update(preds, target, "no")

    def update_bdt(self, preds: torch.Tensor, target: torch.Tensor) -> None:
        """Same statistics straight from (B, D, T) tensors (no materialised rearrange)."""
        self._update(preds, target, "bdt")

    def per_output(self) -> torch.Tensor:
        r, _ = ops.pearson_finalize(_merged_stats(self.stats, self.shift if self._has_shift else None)[0])
        return r

    def compute(self) -> torch.Tensor:
        _, mean = ops.pearson_finalize(_merged_stats(self.stats, self.shift if self._has_shift else None)[0], want_mean=True)
        return mean[0]

    def reset(self) -> None:
        ops.zero_(self.stats) if self.stats.is_cuda else self.stats.zero_()
        self._has_shift = False

    def forward(self, preds, target):
        self.update(preds, target)
        return self.compute()


class GroupedMetric(nn.Module):
    """Per-group (subject) Pearson: one statistics block per group id (metrics/base.py:39-91).  ``compute()`` returns
    ``{str(group_id): float}`` for every group that received samples."""

    MAX_GROUPS = 64

    def __init__(self, metric_name: str = "MultidimPearsonCorrCoef", kwargs: dict[str, tp.Any] | None = None) -> None:
        super().__init__()
        if metric_name != "MultidimPearsonCorrCoef":
            raise NotImplementedError(f"GroupedMetric({metric_name}) is not on the TRIBE path (defaults.py:113-118)")
        self.metric_kwargs = kwargs or {}
        self.num_outputs = int(self.metric_kwargs.get("num_outputs", 1))
        self.register_buffer("stats", torch.zeros(self.MAX_GROUPS, 6, self.num_outputs, dtype=torch.float64), persistent=False)
        self.register_buffer("bad_group", torch.zeros(1, dtype=torch.int32), persistent=False)
        self.register_buffer("shift", torch.zeros(2, self.num_outputs, dtype=torch.float32), persistent=False)  # one pivot per parcel for all groups
        self._has_shift = False

    def _prep(self, preds, groups, n_expected):
        if not preds.is_cuda:
            raise TribeError("GroupedMetric needs CUDA tensors (no CPU fallback)")
        if self.stats.device != preds.device:
            self.stats, self.bad_group, self.shift = self.stats.to(preds.device), self.bad_group.to(preds.device), self.shift.to(preds.device)
        if groups is None:
            groups = torch.zeros(n_expected, dtype=torch.int64, device=preds.device)
        groups = groups.flatten().to(preds.device, torch.int64).contiguous()
        assert len(groups) == n_expected, f"Groups must be the same shape as preds/target, got {groups.shape} and {preds.shape}"
        ops.check_subjects(groups, self.MAX_GROUPS, self.bad_group)
        return groups

    def _update(self, preds, target, groups, layout):
        groups = self._prep(preds, groups, preds.shape[0])
        preds, target = preds.detach().float().contiguous(), target.detach().to(preds.device).float().contiguous()
        if preds.numel() == 0:
            return
        if not self._has_shift:
            ops.pearson_pick_shift(preds, target, self.shift, layout=layout)
            self._has_shift = True
        ops.pearson_stats(preds, target, self.stats, layout=layout, group=groups, n_groups=self.MAX_GROUPS, shift=self.shift)

    def update(self, preds: torch.Tensor, target: torch.Tensor, groups: tp.Optional[torch.Tensor] = None) -> None:
        self._update(preds, target, groups, "no")

    def update_bdt(self, preds: torch.Tensor, target: torch.Tensor, groups: tp.Optional[torch.Tensor] = None) -> None:
        self._update(preds, target, groups, "bdt")

    def compute(self) -> dict[str, float]:
        if int(self.bad_group.item()):
            raise TribeError(f"GroupedMetric saw a group id outside [0, {self.MAX_GROUPS})")
        st = _merged_stats(self.stats, self.shift if self._has_shift else None)
        counts = st[:, 0, 0].cpu()
        out = {}
        for gid in torch.nonzero(counts > 0).flatten().tolist():
            _, mean = ops.pearson_finalize(st[gid], want_mean=True)
            out[str(gid)] = mean.item()
        return out

    def reset(self) -> None:
        ops.zero_(self.stats) if self.stats.is_cuda else self.stats.zero_()
        ops.zero_(self.bad_group) if self.bad_group.is_cuda else self.bad_group.zero_()
        self._has_shift = False

    def __repr__(self) -> str:
        return "GroupedMetric(MultidimPearsonCorrCoef)"


@torch.no_grad()
def compute_multidim_pearson(model: nn.Module, loader: tp.Iterable, parcel_slice: slice | None = None, distributed: str | None = "auto",
                             n_outputs: int | None = None, n_windows: int | None = None, timings: dict | None = None) -> np.ndarray:
    """``Experiment.compute_multidim_pearson`` (main.py:459-477): eval-mode batched predict, then per-parcel Pearson r
    over all (window, TR) rows.  Predictions never leave the device; the 1000-iteration scipy loop becomes one
    statistics kernel per batch + one finalize.  ``parcel_slice`` restricts the statistics to a shard of parcels.

    Rank-aware: under ``torch.distributed`` (world size > 1) every rank passes ITS shard of the windows and every rank
    receives r over ALL ranks' windows.  ``distributed="stats"`` (what "auto" selects) all-reduces the fp64 sufficient
    statistics (6 x O doubles); ``"parcels"`` keeps the predictions, re-lays them to parcel shards with one all-to-all
    and reduces 1000/G parcels per rank (BASELINE config 5, ``parallel.sharded_pearson``); ``None`` = local windows only.
    ``n_windows`` (optional, "parcels" mode): this rank's window count — predictions are then written in place into one
    preallocated buffer by the readout GEMM.  ``timings`` (optional dict) receives CUDA-event pairs per stage."""
    model.eval()
    ws = _world_size()
    mode = ("stats" if ws > 1 else None) if distributed == "auto" else distributed
    if ws == 1 and mode == "stats":
        mode = None
    stats = shift = None
    kept_p, kept_t, buf_p, buf_t, filled = [], [], None, None, 0

    def stamp(name, first=None):
        if timings is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            timings.setdefault(name, []).append(e)

    stamp("predict")
    for batch in loader:
        y_true = batch.data["fmri"]
        if y_true.ndim == 4:
            y_true = y_true.squeeze(-1)
        if mode == "parcels" and n_windows is not None and parcel_slice is None and hasattr(model, "_engine"):
            bsz = y_true.shape[0]
            if buf_p is None:
                dev = model._engine.device if model._engine.device is not None else torch.device("cuda", torch.cuda.current_device())
                buf_p = torch.empty(n_windows, *y_true.shape[1:], device=dev, dtype=torch.float32)
                buf_t = torch.empty_like(buf_p)
            if filled + bsz > n_windows:
                raise TribeError(f"compute_multidim_pearson: more windows than n_windows={n_windows}")
            model(batch, out=buf_p[filled: filled + bsz])
            ops.copy_(buf_t[filled: filled + bsz], y_true.to(buf_t.device, torch.float32).contiguous())
            filled += bsz
            continue
        y_pred = model(batch)
        y_true = y_true.to(y_pred.device, torch.float32)
        if parcel_slice is not None:
            y_pred, y_true = y_pred[:, parcel_slice], y_true.contiguous()[:, parcel_slice]  # read in place by the kernel
        if mode == "parcels":
            kept_p.append(y_pred.float().contiguous()), kept_t.append(y_true.contiguous())
            continue
        if stats is None:
            stats = torch.zeros(1, 6, y_pred.shape[1], device=y_pred.device, dtype=torch.float64)
            shift = torch.empty(2, y_pred.shape[1], device=y_pred.device, dtype=torch.float32)
            ops.pearson_pick_shift(y_pred.float(), y_true, shift, layout="bdt")
        ops.pearson_stats(y_pred.float(), y_true, stats, layout="bdt", shift=shift)
    stamp("predict")
    if mode == "parcels":
        from . import parallel

        if buf_p is not None:
            p_all, t_all = buf_p[:filled], buf_t[:filled]
        else:
            p_all, t_all = torch.cat(kept_p), torch.cat(kept_t)
        r = parallel.sharded_pearson(p_all, t_all, timings=timings)
        return r.cpu().numpy().astype(np.float32)
    if stats is None:  # a rank without windows still has to take part in the reduction
        if mode != "stats":
            raise TribeError("compute_multidim_pearson: empty loader")
        o = n_outputs if n_outputs is not None else getattr(model, "n_outputs", None)
        if o is None:
            raise TribeError("compute_multidim_pearson: a rank without windows needs n_outputs")
        if parcel_slice is not None:
            o = len(range(*parcel_slice.indices(o)))
        stats = torch.zeros(1, 6, o, device=torch.device("cuda", torch.cuda.current_device()), dtype=torch.float64)
    stamp("pearson")
    if mode == "stats":
        stats = _merged_stats(stats, shift)
    r, _ = ops.pearson_finalize(stats[0])
    stamp("pearson")
    return r.cpu().numpy().astype(np.float32)


@torch.no_grad()
def pearson_from_host(preds: torch.Tensor, trues: torch.Tensor, chunk_windows: int = 128, device=None) -> np.ndarray:
    """Per-parcel Pearson r of HOST-resident prediction / target arrays (what main.py:470-473 holds after its predict
    loop): (N_windows, O, T) — or (N_rows, O) row-major — float32, ideally pinned.  Chunks are copied to two alternating
    device slots on a copy stream while the statistics kernel consumes the previous chunk; only r (O floats) returns."""
    if preds.shape != trues.shape or preds.dtype != torch.float32 or trues.dtype != torch.float32:
        raise TribeError("pearson_from_host: preds/trues must be float32 tensors of the same shape")
    if not torch.cuda.is_available():
        raise TribeError("pearson_from_host needs a CUDA device (no CPU fallback)")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    layout = "bdt" if preds.dim() == 3 else "no"
    n, o = preds.shape[0], preds.shape[1]
    if layout == "no":
        chunk_windows *= 100
    stats = torch.zeros(1, 6, o, device=dev, dtype=torch.float64)
    shift = torch.empty(2, o, device=dev, dtype=torch.float32)
    main, copy = torch.cuda.current_stream(dev), torch.cuda.Stream(dev)
    slots = [(torch.empty((min(chunk_windows, n),) + tuple(preds.shape[1:]), device=dev), torch.empty((min(chunk_windows, n),) + tuple(preds.shape[1:]), device=dev))
             for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]   # slot consumed by the statistics kernel
    ready = [torch.cuda.Event() for _ in range(2)]  # slot filled by the copy stream
    for i, lo in enumerate(range(0, n, chunk_windows)):
        hi, s = min(n, lo + chunk_windows), i % 2
        with torch.cuda.stream(copy):
            if i >= 2:
                copy.wait_event(free[s])
            slots[s][0][: hi - lo].copy_(preds[lo:hi], non_blocking=True)
            slots[s][1][: hi - lo].copy_(trues[lo:hi], non_blocking=True)
            ready[s].record(copy)
        main.wait_event(ready[s])
        if i == 0:
            ops.pearson_pick_shift(slots[s][0][: hi - lo], slots[s][1][: hi - lo], shift, layout=layout)
        ops.pearson_stats(slots[s][0][: hi - lo], slots[s][1][: hi - lo], stats, layout=layout, shift=shift)
        free[s].record(main)
    r, _ = ops.pearson_finalize(stats[0])
    return r.cpu().numpy().astype(np.float32)


# ------------------------------------------------------------------------------------------------ retrieval (§8f row 2)
class Rank(nn.Module):
    """Mirror of ``modeling_utils.metrics.metrics.Rank`` (metrics.py:66-168) for the unlabeled use of the TRIBE path
    (``x_labels`` / ``y_labels`` None: the true candidate of query b is row b): rank of the true target among the
    cosine-normalised scores, ties averaged.  State ``ranks`` is concatenated over updates (and over ranks on compute)."""

    is_differentiable = False
    higher_is_better = False

    def __init__(self, reduction: str = "median", relative: bool = False) -> None:
        super().__init__()
        self.reduction = reduction
        self.relative = relative
        self._ranks: list[torch.Tensor] = []

    @property
    def ranks(self) -> torch.Tensor:
        if not self._ranks:
            return torch.zeros(0)
        return torch.cat(self._ranks)

    @torch.no_grad()
    def update(self, x: torch.Tensor, y: torch.Tensor, x_labels=None, y_labels=None) -> None:
        if x_labels is not None or y_labels is not None:
            raise NotImplementedError("labeled retrieval is not on the TRIBE path (pl_module.py:100-101 passes no labels)")
        if not x.is_cuda:
            raise TribeError("retrieval metric needs CUDA tensors (no CPU fallback)")
        assert x.shape[0] == y.shape[0]
        r, _ = ops.retrieval_ranks(x.detach().float().contiguous(), y.detach().to(x.device).float().contiguous())
        if self.relative:
            r = r / y.shape[0]
        self._ranks.append(r)

    @torch.no_grad()
    def update_bdt(self, preds: torch.Tensor, target: torch.Tensor) -> None:
        """``update(preds.mean(-1), target.mean(-1))`` of pl_module.py:100-101 with the time average in our kernel."""
        if not preds.is_cuda:
            raise TribeError("retrieval metric needs CUDA tensors (no CPU fallback)")
        self.update(ops.mean_lastdim(preds.detach().float().contiguous()), ops.mean_lastdim(target.detach().to(preds.device).float().contiguous()))

    def _gathered(self) -> torch.Tensor:
        r = self.ranks
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:  # dist_reduce_fx="cat"
            ws = dist.get_world_size()
            dev = r.device if r.is_cuda else torch.device("cuda", torch.cuda.current_device())
            n = torch.tensor([r.numel()], device=dev)
            sizes = [torch.zeros_like(n) for _ in range(ws)]
            dist.all_gather(sizes, n)
            width = int(max(s.item() for s in sizes))
            mine = torch.zeros(width, device=dev)
            mine[: r.numel()] = r.to(dev)
            parts = [torch.empty_like(mine) for _ in range(ws)]
            dist.all_gather(parts, mine)
            r = torch.cat([p[: int(s.item())] for p, s in zip(parts, sizes)])
        return r

    def compute(self) -> torch.Tensor:
        r = self._gathered()
        if self.reduction == "mean":
            return torch.mean(r)
        if self.reduction == "median":
            return torch.median(r)
        if self.reduction == "std":
            return torch.std(r)
        raise ValueError(f'Unknown aggregation {self.reduction} for computing metric. Available aggregations are: "mean", "median" or "std".')

    def reset(self) -> None:
        self._ranks = []

    def forward(self, x, y):
        self.update(x, y)
        return self.compute()


class TopkAcc(Rank):
    """``modeling_utils.metrics.metrics.TopkAcc`` (metrics.py:194-218): share of queries whose true target is ranked in the
    top k — the ``val/retrieval_top1`` metric of algonauts2025/grids/defaults.py:119-123 (topk=1)."""

    higher_is_better = True

    def __init__(self, topk: int = 5) -> None:
        super().__init__(relative=False)
        self.topk = topk

    def compute(self) -> torch.Tensor:
        return (self._gathered() < self.topk).float().mean()
