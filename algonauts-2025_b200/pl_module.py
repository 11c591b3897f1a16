"""``BrainModule`` — mirror of the reference's LightningModule (algonauts2025/pl_module.py:19-144): same constructor,
``training_step`` / ``validation_step`` / ``test_step`` / ``_run_step`` signatures, same logged keys and metric
dispatch rules, so it drops into algonauts2025/main.py.  Subclasses ``lightning.pytorch.LightningModule`` when
lightning is importable, otherwise a minimal stand-in with the same ``log`` / ``log_dict`` / ``trainer`` surface."""
from __future__ import annotations

import types
import typing as tp
from pathlib import Path

import torch
from torch import nn

from . import losses as L
from . import ops
from .segment import SegmentData

try:  # pragma: no cover - lightning is not in this image
    import lightning.pytorch as pl

    _Base = pl.LightningModule
except Exception:  # noqa: BLE001

    class _Base(nn.Module):
        """Stand-in for ``pl.LightningModule`` (automatic optimisation is driven by trainer.MiniTrainer)."""

        def __init__(self) -> None:
            super().__init__()
            self.logged: dict[str, tp.Any] = {}
            self.trainer = types.SimpleNamespace(estimated_stepping_batches=1000)

        def log(self, name, value, **kwargs):
            # like Lightning's logger connector: keep the value, not its autograd graph (and the workspaces it pins)
            self.logged[name] = value.detach() if torch.is_tensor(value) else value

        def log_dict(self, d, **kwargs):
            for name, value in d.items():
                self.log(name, value)

        def on_validation_epoch_end(self) -> None:
            return None

        def on_test_epoch_end(self) -> None:
            return None


class _FlattenFn(torch.autograd.Function):
    """``rearrange(x, "b d t -> (b t) d")`` (pl_module.py:54-55) as the tiled transpose kernel; its adjoint is the same
    kernel with the two dims swapped."""

    @staticmethod
    def forward(ctx, x):
        ctx.shape = x.shape
        return ops.transpose_last2(x.detach().float().contiguous()).view(-1, x.shape[1])

    @staticmethod
    def backward(ctx, g):
        b, d, t = ctx.shape
        return ops.transpose_last2(g.contiguous().float().view(b, t, d))


def _flatten_bdt(x: torch.Tensor) -> torch.Tensor:
    if x.is_cuda and x.dim() == 3:
        return _FlattenFn.apply(x)
    return x.permute(0, 2, 1).reshape(-1, x.shape[1])  # "b d t -> (b t) d"


class BrainModule(_Base):
    def __init__(self, model: nn.Module, loss: nn.Module, optim_config: tp.Any, metrics: dict[str, tp.Any], max_epochs: int = 100,
                 checkpoint_path: Path | None = None, config: dict[str, tp.Any] | None = None) -> None:
        super().__init__()
        self.model = model
        self.checkpoint_path = checkpoint_path
        self.config = config
        self.optim_config = optim_config
        self.max_epochs = max_epochs
        self.loss = loss
        self.metrics = metrics

    def forward(self, batch):
        return self.model(batch)

    def _run_step(self, batch: SegmentData, batch_idx, step_name):
        y_pred = self.forward(batch)  # B, D, T  (CUDA)
        y_true = batch.data["fmri"].to(y_pred.device, non_blocking=True)  # B, D, T
        if step_name == "val":
            y_true = y_true[:, :, 0:]
            y_pred = y_pred[:, :, 0:]
        # the grid's losses run fused on the (B, D, T) tensors (their value is invariant to the (b t) d rearrange, or
        # — PearsonLoss — the kernel indexes parcels in place); anything else gets the flattened matrices
        loss = L.fused_loss(self.loss, y_pred, y_true)
        if loss is None:
            y_pred_flat, y_true_flat = _flatten_bdt(y_pred), _flatten_bdt(y_true)
            loss = self.loss(y_pred_flat, y_true_flat)

        if hasattr(self.model, "compute_contrastive_loss"):
            contrastive_losses = self.model.compute_contrastive_loss(batch)
            if contrastive_losses:
                weight = getattr(self.model.config, "contrastive_weight", 0.0)
                total_contrastive = 0.0
                for name, c_loss in contrastive_losses.items():
                    self.log(f"{step_name}/contrastive/{name}", c_loss, on_step=False, on_epoch=True, logger=True, prog_bar=False,
                             batch_size=y_pred.shape[0])
                    total_contrastive = total_contrastive + c_loss
                total_contrastive = total_contrastive / max(1, len(contrastive_losses))
                loss = loss + weight * total_contrastive
        log_kwargs = {"on_step": step_name == "train", "on_epoch": True, "logger": True, "prog_bar": True, "batch_size": y_pred.shape[0]}
        self.log(f"{step_name}/loss", loss, **log_kwargs)

        for metric_name, metric in self.metrics.items():
            if metric_name.startswith(step_name):
                yp, yt = y_pred.detach(), y_true
                if "grouped" in metric.__class__.__name__.lower():
                    if hasattr(metric, "update_bdt"):
                        metric.update_bdt(yp, yt, groups=batch.data["subject_id"])
                    else:
                        groups = batch.data["subject_id"].to(yp.device).repeat_interleave(yp.shape[2], 0)
                        metric.update(_flatten_bdt(yp), _flatten_bdt(yt), groups=groups)
                else:
                    if "retrieval" in metric_name:
                        if hasattr(metric, "update_bdt"):
                            metric.update_bdt(yp, yt)  # the time average is fused into the metric's kernels
                        else:
                            metric.update(yp.mean(dim=-1), yt.mean(dim=-1))
                    elif hasattr(metric, "update_bdt"):
                        metric.update_bdt(yp, yt)
                    else:
                        metric.update(_flatten_bdt(yp), _flatten_bdt(yt))
                    self.log(metric_name, metric, **log_kwargs)
        if step_name == "train":
            # training_step discards the predictions; skip the reference's per-step D2H copy + stream sync
            return loss, y_pred.detach(), y_true
        return loss, y_pred.detach().cpu(), y_true.detach().cpu()

    def on_val_or_test_epoch_end(self, step_name: str) -> None:
        for metric_name, metric in self.metrics.items():
            if metric_name.startswith(step_name):
                if "grouped" in metric.__class__.__name__.lower():
                    metric_dict = {metric_name + "/" + k: v for k, v in metric.compute().items()}
                    self.log_dict(metric_dict)

    def on_validation_epoch_end(self) -> None:
        self.on_val_or_test_epoch_end("val")
        return super().on_validation_epoch_end()

    def on_test_epoch_end(self) -> None:
        self.on_val_or_test_epoch_end("test")
        return super().on_test_epoch_end()

    def training_step(self, batch: SegmentData, batch_idx):
        loss, _, _ = self._run_step(batch, batch_idx, step_name="train")
        return loss

    def validation_step(self, batch: SegmentData, batch_idx):
        _, y_pred, y_true = self._run_step(batch, batch_idx, step_name="val")
        return y_pred, y_true

    def test_step(self, batch: SegmentData, batch_idx):
        _, y_pred, y_true = self._run_step(batch, batch_idx, step_name="test")
        return y_pred, y_true

    def configure_optimizers(self):
        optim_config = self.optim_config.copy()
        unfrozen_params = [p for p in self.parameters() if p.requires_grad]
        out = optim_config.build(unfrozen_params, total_steps=self.trainer.estimated_stepping_batches)
        # A stock torch.optim.Adam (the reference recipe, defaults.py:126-141) is adopted in place by the fused
        # Adam(+bf16 shadow) kernel; any other optimizer is left untouched.
        from .optim import TribeAdam

        opt = out["optimizer"] if isinstance(out, dict) else out
        TribeAdam.adopt(opt, self.model)
        return out
