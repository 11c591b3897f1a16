"""Stochastic weight averaging over the flat parameter buffer — the reference enables Lightning's
``StochasticWeightAveraging(swa_lrs=1e-5, swa_epoch_start=0.6, annealing_epochs=int(0.4 * n_epochs))``
(algonauts2025/main.py:365-373), whose per-epoch ``update_parameters`` is torch's ``AveragedModel`` rule
``avg += (p - avg) / (n_averaged + 1)`` applied tensor by tensor.  Here all 941.5 M parameters live in ONE flat fp32
buffer (``engine.FlatParams``), so an update is a single 12 B/parameter streaming kernel."""
from __future__ import annotations

import torch

from . import ops
from ._lib import TribeError


class SwaAverager:
    def __init__(self, model):
        self.model = model
        self.n_averaged = 0
        self.avg: torch.Tensor | None = None

    def _flat(self):
        eng = getattr(self.model, "_engine", None)
        if eng is None:
            raise TribeError("SwaAverager needs an FmriEncoder (flat parameter buffer)")
        eng._check_flat()
        return eng.flat

    @torch.no_grad()
    def update_parameters(self) -> None:
        flat = self._flat()
        if self.avg is None:
            self.avg = torch.zeros_like(flat.flat)
        ops.swa_update(self.avg, flat.flat, self.n_averaged)
        self.n_averaged += 1

    def averaged_state_dict(self) -> dict:
        """state_dict-shaped views of the averaged weights (buffers are taken from the live model)."""
        flat = self._flat()
        if self.avg is None:
            raise TribeError("no parameters averaged yet")
        sd = {k: v for k, v in self.model.state_dict().items()}
        for name, p in flat.params.items():
            off = flat.offsets[name]
            sd[name] = self.avg[off: off + p.numel()].view(p.shape)
        return sd

    @torch.no_grad()
    def swap_into_model(self) -> None:
        """Exchange live and averaged weights in place (what Lightning does at the end of fit / around validation)."""
        flat = self._flat()
        if self.avg is None:
            raise TribeError("no parameters averaged yet")
        tmp = flat.flat.clone()
        flat.flat.copy_(self.avg)
        self.avg.copy_(tmp)
        flat._sig = None  # parameters are views of the flat buffer; the bf16 shadow is re-cast at the next forward
