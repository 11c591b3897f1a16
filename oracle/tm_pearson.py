"""TEST INFRASTRUCTURE — CPU restatement of ``torchmetrics.PearsonCorrCoef(num_outputs=O)``.

The reference's validation metric is ``MultidimPearsonCorrCoef`` = ``torchmetrics.PearsonCorrCoef`` + ``.mean()``
(reference ``modeling_utils/modeling_utils/metrics/base.py:26-29``), updated from
``algonauts2025/pl_module.py:93-106``.  ``torchmetrics>=1.1.2`` (``modeling_utils/pyproject.toml:11``) is not installed
and not vendored; the streaming update below restates its published algorithm (state names are the ones the reference
itself lists at ``metrics/metrics.py:38-44``).  **PARITY UNPINNED** by any reference test; the batch-Pearson value it
converges to is pinned against ``scipy.stats.pearsonr`` (the reference's own final-eval call, ``main.py:474-476``) in
``tests/test_oracle.py``.
"""
from __future__ import annotations

import torch
from torch import nn


class Metric(nn.Module):
    """Minimal ``torchmetrics.Metric``: ``add_state`` registers resettable buffers."""

    def __init__(self, **kwargs):
        super().__init__()
        self._defaults = {}

    def add_state(self, name, default, dist_reduce_fx=None, persistent=False):
        self._defaults[name] = default.clone() if isinstance(default, torch.Tensor) else list(default)
        if isinstance(default, torch.Tensor):
            self.register_buffer(name, default.clone(), persistent=persistent)
        else:
            setattr(self, name, list(default))

    def reset(self):
        for name, default in self._defaults.items():
            if isinstance(default, torch.Tensor):
                getattr(self, name).copy_(default.to(getattr(self, name).device))
            else:
                setattr(self, name, list(default))

    @property
    def device(self):
        for b in self.buffers():
            return b.device
        return torch.device("cpu")

    def forward(self, *a, **k):
        self.update(*a, **k)
        return self.compute()


class PearsonCorrCoef(Metric):
    def __init__(self, num_outputs: int = 1, **kwargs):
        super().__init__(**kwargs)
        self.num_outputs = num_outputs
        for name in ("mean_x", "mean_y", "var_x", "var_y", "corr_xy", "n_total"):
            self.add_state(name, torch.zeros(num_outputs), dist_reduce_fx=None)

    def update(self, preds: torch.Tensor, target: torch.Tensor) -> None:
        if preds.ndim == 1:
            preds, target = preds[:, None], target[:, None]
        n = preds.shape[0]
        cond = bool(self.n_total.mean() > 0) or n == 1
        if cond:
            mx_new = (self.n_total * self.mean_x + preds.sum(0)) / (self.n_total + n)
            my_new = (self.n_total * self.mean_y + target.sum(0)) / (self.n_total + n)
        else:
            mx_new, my_new = preds.mean(0), target.mean(0)
        self.n_total += n
        if cond:
            self.var_x += ((preds - mx_new) * (preds - self.mean_x)).sum(0)
            self.var_y += ((target - my_new) * (target - self.mean_y)).sum(0)
        else:
            self.var_x += preds.var(0) * (n - 1)
            self.var_y += target.var(0) * (n - 1)
        self.corr_xy += ((preds - mx_new) * (target - self.mean_y)).sum(0)
        self.mean_x.copy_(mx_new)
        self.mean_y.copy_(my_new)

    def compute(self) -> torch.Tensor:
        nb = self.n_total
        var_x, var_y, corr_xy = self.var_x / (nb - 1), self.var_y / (nb - 1), self.corr_xy / (nb - 1)
        r = (corr_xy / (var_x * var_y).sqrt()).clamp(-1.0, 1.0)
        return r.squeeze()
