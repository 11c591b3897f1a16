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


def _is_torchmetrics(metric) -> bool:
    """Lightning's ``self.log(name, metric_object)`` accepts ``torchmetrics.Metric`` instances only (it computes / resets
    them at epoch end); anything else must be logged by value."""
    try:
        import torchmetrics
    except Exception:  # noqa: BLE001
        return False
    return isinstance(metric, torchmetrics.Metric)


def _flatten_bdt(x: torch.Tensor) -> torch.Tensor:
    if x.is_cuda and x.dim() == 3:
        return _FlattenFn.apply(x)
    return x.permute(0, 2, 1).reshape(-1, x.shape[1])  # "b d t -> (b t) d"


class BrainModule(_Base):
    """Public surface = the reference's: ``__init__(model, loss, optim_config, metrics, max_epochs, checkpoint_path,
    config)``, ``forward``, ``_run_step(batch, batch_idx, step_name) -> (loss, y_pred, y_true)``, ``training_step`` /
    ``validation_step`` / ``test_step``, the two epoch-end hooks and ``configure_optimizers``.  The step itself is split
    into loss / contrastive / metric stages that all work on the (B, D, T) device tensors."""

    GROUPED_TAG, RETRIEVAL_TAG = "grouped", "retrieval"  # dispatch rules of pl_module.py:93-106

    def __init__(self, model: nn.Module, loss: nn.Module, optim_config: tp.Any, metrics: dict[str, tp.Any], max_epochs: int = 100,
                 checkpoint_path: Path | None = None, config: dict[str, tp.Any] | None = None) -> None:
        super().__init__()
        self.model, self.loss, self.metrics = model, loss, metrics
        self.optim_config, self.max_epochs = optim_config, max_epochs
        self.checkpoint_path, self.config = checkpoint_path, config
        self._grad_sync = None  # parallel.data_parallel(...) installed by configure_optimizers in multi-rank jobs

    def forward(self, batch):
        return self.model(batch)

    # ------------------------------------------------------------------------------------------------ stages of a step
    def _primary_loss(self, pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """The configured loss on the step's predictions.  The grid's losses (run_ensemble.py:29) have fused kernels that
        read the (B, D, T) tensors directly — their value does not depend on the ``(b t) d`` rearrange of
        pl_module.py:54-55, or (PearsonLoss) the kernel indexes parcels in place; any other module receives the
        flattened matrices exactly like in the reference."""
        fused = L.fused_loss(self.loss, pred, target)
        if fused is not None:
            return fused
        return self.loss(_flatten_bdt(pred), _flatten_bdt(target))

    def _contrastive_term(self, batch, stage: str, n_windows: int):
        """``contrastive_weight * mean(per-modality InfoNCE)`` (pl_module.py:59-77), or None when the model has no
        contrastive branch / it is disabled.  Each modality's loss is logged under ``{stage}/contrastive/{modality}``."""
        compute = getattr(self.model, "compute_contrastive_loss", None)
        per_modality = compute(batch) if compute is not None else None
        if not per_modality:
            return None
        acc = None
        for modality, value in per_modality.items():
            self.log(f"{stage}/contrastive/{modality}", value, on_step=False, on_epoch=True, logger=True, prog_bar=False, batch_size=n_windows)
            acc = value if acc is None else acc + value
        return getattr(self.model.config, "contrastive_weight", 0.0) * (acc / max(1, len(per_modality)))

    def _feed_metrics(self, stage: str, pred: torch.Tensor, target: torch.Tensor, subject_id, log_opts: dict) -> None:
        for key, metric in self.metrics.items():
            if not key.startswith(stage):
                continue
            on_device = hasattr(metric, "update_bdt")  # our metric classes read (B, D, T) in place
            if self.GROUPED_TAG in type(metric).__name__.lower():
                if on_device:
                    metric.update_bdt(pred, target, groups=subject_id)
                else:
                    per_row = subject_id.to(pred.device).repeat_interleave(pred.shape[2], 0)
                    metric.update(_flatten_bdt(pred), _flatten_bdt(target), groups=per_row)
                continue  # grouped metrics are logged at epoch end only
            if self.RETRIEVAL_TAG in key and not on_device:
                metric.update(pred.mean(dim=-1), target.mean(dim=-1))
            elif on_device:
                metric.update_bdt(pred, target)  # retrieval metrics fuse the time average into their kernels
            else:
                metric.update(_flatten_bdt(pred), _flatten_bdt(target))
            if _is_torchmetrics(metric):
                self.log(key, metric, **log_opts)  # Lightning computes and resets it at epoch end
            # our kernel-backed metric classes are plain modules: their VALUE is logged at epoch end
            # (on_val_or_test_epoch_end) and they are reset at the start of the next epoch

    def _run_step(self, batch: SegmentData, batch_idx, step_name):
        pred = self.forward(batch)                                              # (B, D, T) on the device
        target = batch.data["fmri"].to(pred.device, non_blocking=True)
        if step_name == "val":                                                  # pl_module.py:50-52 (a no-op slice)
            pred, target = pred[:, :, 0:], target[:, :, 0:]
        n_windows = pred.shape[0]
        loss = self._primary_loss(pred, target)
        extra = self._contrastive_term(batch, step_name, n_windows)
        if extra is not None:
            loss = loss + extra
        log_opts = dict(on_step=step_name == "train", on_epoch=True, logger=True, prog_bar=True, batch_size=n_windows)
        self.log(f"{step_name}/loss", loss, **log_opts)
        self._feed_metrics(step_name, pred.detach(), target, batch.data.get("subject_id"), log_opts)
        if step_name == "train":
            # training_step throws the predictions away: no per-step D2H copy + stream sync as in pl_module.py:107
            return loss, pred.detach(), target
        return loss, pred.detach().cpu(), target.detach().cpu()

    # ------------------------------------------------------------------------------------------------ Lightning hooks
    def training_step(self, batch: SegmentData, batch_idx):
        if self._grad_sync is not None:
            self._grad_sync.begin_step()
        return self._run_step(batch, batch_idx, step_name="train")[0]

    def on_before_optimizer_step(self, optimizer, *args, **kwargs) -> None:
        """Lightning calls this after ``loss.backward()`` and right before ``optimizer.step()``: the data-parallel step tail
        (gradient reduction, rank-sharded Adam, shadow multicast — or the NCCL all-reduce fallback) closes here.  Under a
        Lightning DDP strategy torch's own reducer sees no gradient (they are written by hand) and stays a no-op."""
        if self._grad_sync is not None:
            self._grad_sync.finish_step()

    def validation_step(self, batch: SegmentData, batch_idx):
        return self._run_step(batch, batch_idx, step_name="val")[1:]

    def test_step(self, batch: SegmentData, batch_idx):
        return self._run_step(batch, batch_idx, step_name="test")[1:]

    def on_val_or_test_epoch_end(self, step_name: str) -> None:
        for key, metric in self.metrics.items():
            if not key.startswith(step_name):
                continue
            if self.GROUPED_TAG in type(metric).__name__.lower():
                self.log_dict({f"{key}/{group}": value for group, value in metric.compute().items()})
            elif not _is_torchmetrics(metric):
                self.log(key, metric.compute())  # epoch-level value (what on_epoch=True logging of a Metric object yields)

    def _reset_metrics(self, step_name: str) -> None:
        for key, metric in self.metrics.items():
            if key.startswith(step_name) and not _is_torchmetrics(metric) and hasattr(metric, "reset"):
                metric.reset()  # Lightning resets torchmetrics objects itself; plain modules would accumulate across epochs

    def on_validation_epoch_start(self) -> None:
        self._reset_metrics("val")

    def on_test_epoch_start(self) -> None:
        self._reset_metrics("test")

    def on_validation_epoch_end(self) -> None:
        self.on_val_or_test_epoch_end("val")
        return super().on_validation_epoch_end()

    def on_test_epoch_end(self) -> None:
        self.on_val_or_test_epoch_end("test")
        return super().on_test_epoch_end()

    def configure_optimizers(self):
        trainable = [p for p in self.parameters() if p.requires_grad]
        built = self.optim_config.copy().build(trainable, total_steps=self.trainer.estimated_stepping_batches)
        # a stock torch.optim.Adam (the reference recipe, defaults.py:126-141) is adopted in place by the fused
        # Adam(+bf16 shadow) kernel; any other optimizer is left untouched
        from .optim import TribeAdam

        optimizer = built["optimizer"] if isinstance(built, dict) else built
        TribeAdam.adopt(optimizer, self.model)
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and not getattr(self.model, "independent_replica", False) \
                and hasattr(self.model, "_engine"):
            from . import parallel

            # multi-rank job (Lightning DDP, main.py:388-394): torch DDP cannot see this model's gradients — install our own
            # gradient path (collective: every rank reaches configure_optimizers)
            self._grad_sync = parallel.data_parallel(self.model, optimizer)
        return built
