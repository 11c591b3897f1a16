"""Minimal stand-in for the slice of ``lightning.pytorch.Trainer`` the hot path runs under (lightning is not
installed here): Lightning's *automatic optimisation* for one batch is
``zero_grad(set_to_none=True) -> training_step -> loss.backward() -> optimizer.step() -> scheduler.step()`` with no
gradient clipping / accumulation and precision 32-true (reference: algonauts2025/main.py:388-404)."""
from __future__ import annotations

import os
import typing as tp

import torch


def default_optimizer(params, total_steps: int, lr: float = 1e-4, model=None, fused_kernel: bool = True):
    """Adam(lr=1e-4, weight_decay=0) + OneCycleLR(max_lr=1e-4, pct_start=0.1), stepped per batch
    (algonauts2025/grids/defaults.py:126-141, modeling_utils/optimizers/base.py:84-96).  With ``model`` given the
    stock torch.optim.Adam instance is adopted by ``TribeAdam`` (same object, fused Adam + bf16 shadow kernel) — what
    ``BrainModule.configure_optimizers`` does; otherwise torch's own fused multi-tensor Adam runs."""
    params = list(params)
    use_kernel = model is not None and fused_kernel
    opt = torch.optim.Adam(params, lr=lr, weight_decay=0.0, fused=params[0].is_cuda and not use_kernel)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=lr, pct_start=0.1, total_steps=max(total_steps, 2))
    if use_kernel:
        from .optim import TribeAdam

        TribeAdam.adopt(opt, model)
    return opt, sched


class MiniTrainer:
    """``use_graphs=True``: steady-state steps are replayed as whole-step CUDA graphs (``graphed.GraphedTrainStep``);
    results are identical to the eager path (same kernels, same RNG draws), only the host enqueue cost disappears.
    ``graph_collectives``: also capture the NCCL gradient all-reduces of ``grad_sync`` (data-parallel runs).
    ``overlap_optimizer``: run the optimizer layer by layer behind the backward pass (``parallel.StepOverlap``)."""

    def __init__(self, module, optimizer, scheduler=None, grad_sync=None, use_graphs: bool = False, graph_collectives: bool = True,
                 overlap_optimizer: bool = False, fuse_optimizer: bool | None = None):
        self.module, self.optimizer, self.scheduler, self.grad_sync = module, optimizer, scheduler, grad_sync
        self.global_step = 0
        self.graph_collectives = graph_collectives
        self._graphed = None
        model = getattr(module, "model", None)
        # optimizer-in-backward (optim.TribeAdam.fuse_backward): single-GPU / ensemble-member steps have nothing between
        # the weight gradients and Adam, so the update can ride in the wgrad GEMM epilogues.  Measured on B200 this is a
        # wash (profiles/r02_adam_in_wgrad_ab.txt: the wgrad GEMMs already use ~80 % of an SM's L2 read bandwidth, the
        # state traffic lengthens them by what the separate pass cost), so it is opt-in (or TRIBE_FUSED_ADAM=1).
        if fuse_optimizer is None:
            fuse_optimizer = os.environ.get("TRIBE_FUSED_ADAM", "0") == "1"
        self.fuse_optimizer = bool(fuse_optimizer) and grad_sync is None and not overlap_optimizer and hasattr(optimizer, "arm_fused_backward")
        if hasattr(optimizer, "arm_fused_backward"):
            optimizer.fuse_backward = self.fuse_optimizer
        if overlap_optimizer and model is not None and hasattr(optimizer, "step_bucket"):
            # each encoder layer's Adam step runs on a side stream as soon as that layer's gradients are final
            # (after their all-reduce in data-parallel runs), hidden behind the backward of the earlier layers
            from .parallel import StepOverlap

            if self.grad_sync is None:
                self.grad_sync = StepOverlap(model, optimizer=optimizer)
            else:
                self.grad_sync.optimizer = optimizer
        if model is not None and hasattr(model, "defer_subject_check"):
            model.defer_subject_check = True  # no host sync inside the step; raised one call later (see model.py)
        if use_graphs:
            from .graphed import GraphedTrainStep

            self._graphed = GraphedTrainStep(self)

    def run_step_body(self, batch) -> torch.Tensor:
        """Device work of one automatic-optimisation step (everything a CUDA graph may capture)."""
        self.optimizer.zero_grad(set_to_none=True)
        if self.fuse_optimizer:
            self.optimizer.arm_fused_backward()
        if self.grad_sync is not None:
            self.grad_sync.begin_step()
        loss = self.module.training_step(batch, self.global_step)
        loss.backward()
        if self.grad_sync is not None:
            self.grad_sync.finish_step()
        self.optimizer.step()
        return loss.detach()

    def eager_step(self, batch) -> torch.Tensor:
        self.module.train()
        loss = self.run_step_body(batch)
        if self.scheduler is not None:
            self.scheduler.step()
        self.global_step += 1
        return loss

    def train_step(self, batch) -> torch.Tensor:
        if self._graphed is not None:
            return self._graphed.step(batch)
        return self.eager_step(batch)

    @torch.no_grad()
    def validate(self, batches: tp.Iterable) -> dict:
        self.module.eval()
        for name, metric in self.module.metrics.items():
            if name.startswith("val"):
                metric.reset()
        for i, batch in enumerate(batches):
            self.module.validation_step(batch, i)
        if hasattr(getattr(self.module, "model", None), "flush_subject_check"):
            self.module.model.flush_subject_check()
        self.module.on_validation_epoch_end()
        out = dict(getattr(self.module, "logged", {}))
        for name, metric in self.module.metrics.items():
            if name.startswith("val") and "grouped" not in metric.__class__.__name__.lower():
                out[name] = metric.compute()
        return out
