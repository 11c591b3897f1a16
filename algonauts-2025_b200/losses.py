"""Losses of the hot path.  ``nn.MSELoss`` (reference default, algonauts2025/grids/defaults.py:125, applied at
algonauts2025/pl_module.py:56) runs as ONE fused forward+gradient reduction kernel; the other losses the reference's
ensemble grid samples (PearsonLoss / SmoothL1Loss / HuberLoss, algonauts2025/grids/run_ensemble.py:29) likewise."""
from __future__ import annotations

import torch
from torch import nn

from . import ops


class _MseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        loss, grad = ops.mse_fwd_bwd(pred.detach().contiguous(), target.contiguous(), want_grad=pred.requires_grad)
        ctx.save_for_backward(grad) if grad is not None else None
        ctx.has_grad = grad is not None
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        if not ctx.has_grad:
            return None, None
        (grad,) = ctx.saved_tensors
        return ops.scale_dev(grad, g), None


def mse_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """mean((pred - target)^2) over all elements (layout independent, so the (b t) d rearrange is not materialised)."""
    return _MseFn.apply(pred, target.to(pred.device, torch.float32))


def is_plain_mse(loss: nn.Module) -> bool:
    return type(loss) is nn.MSELoss and loss.reduction == "mean"


# ------------------------------------------------------------------------------------------------ grid losses (§8f row 1)
class _PointFn(torch.autograd.Function):
    """SmoothL1 / Huber / L1 (reduction="mean") as ONE fused value + gradient pass (like ``_MseFn``)."""

    @staticmethod
    def forward(ctx, pred, target, kind, param):
        loss, grad = ops.point_loss_fwd_bwd(pred.detach().contiguous(), target.contiguous(), kind, param, want_grad=pred.requires_grad)
        ctx.save_for_backward(grad) if grad is not None else None
        ctx.has_grad = grad is not None
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        if not ctx.has_grad:
            return None, None, None, None
        (grad,) = ctx.saved_tensors
        return ops.scale_dev(grad, g), None, None, None


class _PearsonFn(torch.autograd.Function):
    """``PearsonLoss(dim=1)`` (modeling_utils/losses/losses.py:11-42): forward = per-parcel sufficient statistics
    (the evaluation kernel) + finalize; backward = one element-wise pass using 4 coefficients per parcel."""

    @staticmethod
    def forward(ctx, pred, target, layout, reduction_mean):
        p, t = pred.detach().contiguous(), target.detach().contiguous()
        loss, coef = ops.pearson_loss_fwd(p, t, layout=layout, reduction_mean=reduction_mean, want_coef=pred.requires_grad)
        ctx.layout, ctx.reduction_mean, ctx.has_grad = layout, reduction_mean, coef is not None
        if coef is not None:
            ctx.save_for_backward(p, t, coef)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        if not ctx.has_grad:
            return None, None, None, None
        p, t, coef = ctx.saved_tensors
        up = g.detach().float().reshape(1).contiguous()
        return ops.pearson_loss_bwd(p, t, coef, up, layout=ctx.layout, reduction_mean=ctx.reduction_mean), None, None, None


class PearsonLoss(nn.Module):
    """Mirror of ``modeling_utils.losses.losses.PearsonLoss`` (same constructor / semantics) on the fused kernels.
    ``forward(x, y)`` takes the flattened ``(b t) d`` matrices of pl_module.py:54-56 (dim=1: one correlation per column)."""

    def __init__(self, reduction: str = "mean", dim: int = 1):
        super().__init__()
        self.reduction = reduction
        self.dim = dim

    def forward(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        if self.reduction not in ("mean", "sum"):
            raise ValueError(f"Invalid reduction: {self.reduction}")
        if not x.is_cuda:
            raise ops.TribeError("PearsonLoss needs CUDA tensors (no CPU fallback)")
        if x.dim() != 2 or self.dim not in (0, 1):
            raise NotImplementedError("PearsonLoss kernel path covers 2-D inputs (the TRIBE training path)")
        if self.dim == 0:  # correlate rows instead of columns
            x, y = x.t(), y.t()
        return _PearsonFn.apply(x.float(), y.to(x.device, torch.float32), "no", self.reduction == "mean")


def fused_loss(loss: nn.Module, y_pred: torch.Tensor, y_true: torch.Tensor) -> torch.Tensor | None:
    """The loss of ``_run_step`` computed straight on the (B, D, T) tensors when ``loss`` is one of the grid's losses
    (run_ensemble.py:29) in a layout-independent configuration; None -> the caller flattens and calls the module."""
    y_true = y_true.to(y_pred.device, torch.float32)
    cls = type(loss)
    if is_plain_mse(loss):
        return _MseFn.apply(y_pred, y_true)
    if cls is nn.SmoothL1Loss and loss.reduction == "mean":
        return _PointFn.apply(y_pred, y_true, ops.LOSS_SMOOTH_L1, float(loss.beta))
    if cls is nn.HuberLoss and loss.reduction == "mean":
        return _PointFn.apply(y_pred, y_true, ops.LOSS_HUBER, float(loss.delta))
    if cls is nn.L1Loss and loss.reduction == "mean":
        return _PointFn.apply(y_pred, y_true, ops.LOSS_L1, 0.0)
    if cls.__name__ == "PearsonLoss" and getattr(loss, "dim", None) == 1 and getattr(loss, "reduction", None) in ("mean", "sum") and y_pred.dim() == 3:
        return _PearsonFn.apply(y_pred, y_true, "bdt", loss.reduction == "mean")
    return None
