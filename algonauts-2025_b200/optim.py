"""``TribeAdam`` — ``torch.optim.Adam`` whose ``step()`` runs ONE fused kernel per contiguous parameter run of the flat
parameter buffer and writes the bf16 shadow weights in the same pass (``csrc/optim.cu``).

It IS a ``torch.optim.Adam`` (same constructor, ``param_groups``, ``state`` / ``state_dict`` keys ``step``, ``exp_avg``,
``exp_avg_sq``, hooks, LR-scheduler compatibility incl. OneCycleLR's per-step ``lr`` and ``betas`` updates); a stock
instance built by the reference's optimizer config (modeling_utils/optimizers/base.py:47-48) can be *adopted* in
place, so schedulers that already hold a reference keep working.  Anything the fused path does not cover (amsgrad,
maximize, capturable, foreign parameters) falls through to the stock implementation."""
from __future__ import annotations

import ctypes
import math

import torch

from . import _lib, engine as _engine
from ._lib import check


class TribeAdam(torch.optim.Adam):
    @classmethod
    def adopt(cls, optimizer: torch.optim.Optimizer, model) -> torch.optim.Optimizer:
        """Turn a stock ``torch.optim.Adam`` instance into a TribeAdam in place (no-op for other optimizers)."""
        if type(optimizer) is torch.optim.Adam:
            optimizer.__class__ = cls
            optimizer._tribe_model = model
            # torch hooks `step` per class (profile / pre / post hooks) and LR schedulers additionally wrap the
            # *instance's* bound step captured at their construction: re-establish both for the new class.
            inst_step = optimizer.__dict__.pop("step", None)
            optimizer._patch_step_function()
            if inst_step is not None and getattr(inst_step, "_wrapped_by_lr_sched", False):
                import weakref

                ref, func = weakref.ref(optimizer), cls.step

                def wrapper(*args, **kwargs):
                    opt = ref()
                    opt._opt_called = True
                    return func.__get__(opt, opt.__class__)(*args, **kwargs)

                wrapper._wrapped_by_lr_sched = True
                optimizer.step = wrapper
        return optimizer

    def _flat(self):
        model = getattr(self, "_tribe_model", None)
        eng = getattr(model, "_engine", None) if model is not None else None
        if eng is None or not torch.cuda.is_available():
            return None
        eng._check_flat()
        return eng.flat

    def _fusable(self, flat) -> bool:
        if flat is None:
            return False
        own = {id(p) for p in flat.params.values()}
        for group in self.param_groups:
            if group.get("amsgrad") or group.get("maximize") or group.get("capturable") or group.get("differentiable"):
                return False
            if any(id(p) not in own for p in group["params"]):
                return False
        return True

    @torch.no_grad()
    def step(self, closure=None):
        flat = self._flat()
        if not self._fusable(flat):
            return super().step(closure)
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if flat.bf16 is None:
            flat.refresh_bf16()
        if getattr(flat, "adam_m", None) is None:
            flat.adam_m, flat.adam_v = torch.zeros_like(flat.flat), torch.zeros_like(flat.flat)
        name_of = {id(p): n for n, p in flat.params.items()}
        lib = _lib.load()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for group in self.param_groups:
            lr = float(group["lr"])
            beta1, beta2 = (float(b) for b in group["betas"])
            eps, wd = float(group["eps"]), float(group["weight_decay"])
            todo = []
            for p in group["params"]:
                if p.grad is None:
                    continue
                name = name_of[id(p)]
                off = flat.offsets[name]
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = flat.adam_m[off: off + p.numel()].view(p.shape)
                    st["exp_avg_sq"] = flat.adam_v[off: off + p.numel()].view(p.shape)
                    st["exp_avg"].zero_(), st["exp_avg_sq"].zero_()
                elif st["exp_avg"].data_ptr() != flat.adam_m.data_ptr() + 4 * off:  # state loaded from a checkpoint
                    flat.adam_m[off: off + p.numel()].view(p.shape).copy_(st["exp_avg"])
                    flat.adam_v[off: off + p.numel()].view(p.shape).copy_(st["exp_avg_sq"])
                    st["exp_avg"] = flat.adam_m[off: off + p.numel()].view(p.shape)
                    st["exp_avg_sq"] = flat.adam_v[off: off + p.numel()].view(p.shape)
                if p.grad.data_ptr() != flat.grad.data_ptr() + 4 * off:  # foreign gradient tensor: bring it into the flat buffer
                    flat.gview(name).copy_(p.grad)
                st["step"] += 1
                todo.append((off, off + _engine._round_up(p.numel(), _engine.ALIGN), int(st["step"].item())))
            # merge adjacent parameters with the same step count into one launch
            todo.sort()
            runs = []
            for lo, hi, k in todo:
                if runs and runs[-1][1] == lo and runs[-1][2] == k:
                    runs[-1][1] = hi
                else:
                    runs.append([lo, hi, k])
            for lo, hi, k in runs:
                n = hi - lo
                check(lib.tribe_adam_step(ctypes.c_void_p(flat.flat.data_ptr() + 4 * lo), ctypes.c_void_p(flat.grad.data_ptr() + 4 * lo),
                                          ctypes.c_void_p(flat.adam_m.data_ptr() + 4 * lo), ctypes.c_void_p(flat.adam_v.data_ptr() + 4 * lo),
                                          ctypes.c_void_p(flat.bf16.data_ptr() + 2 * lo), n, lr, beta1, beta2, eps, wd, k, stream), "tribe_adam_step")
        # the shadow weights are already current for the state the post-step hook is about to announce
        flat._sig = (sum(p._version for p in flat.params.values()), flat.opt_steps + 1)
        return loss
