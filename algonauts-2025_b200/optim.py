"""``TribeAdam`` — ``torch.optim.Adam`` whose ``step()`` runs ONE fused kernel per contiguous parameter run of the flat
parameter buffer and writes the bf16 shadow weights in the same pass (``csrc/optim.cu``).

It IS a ``torch.optim.Adam`` (same constructor, ``param_groups``, ``state`` / ``state_dict`` keys ``step``, ``exp_avg``,
``exp_avg_sq``, hooks, LR-scheduler compatibility incl. OneCycleLR's per-step ``lr`` and ``betas`` updates); a stock
instance built by the reference's optimizer config (modeling_utils/optimizers/base.py:47-48) can be *adopted* in
place, so schedulers that already hold a reference keep working.  Anything the fused path does not cover (amsgrad,
maximize, capturable, foreign parameters) falls through to the stock implementation."""
from __future__ import annotations

import ctypes

import torch

from . import _lib, engine as _engine
from ._lib import check


class TribeAdam(torch.optim.Adam):
    _tribe_sharded = None  # parallel.ShardedStep when the step tail runs rank-sharded over NVLink
    # Optimizer-in-backward for the GEMM weights (96 % of the parameters): with ``fuse_backward = True`` an *armed* step
    # (``arm_fused_backward()``, or ``step(closure)`` as Lightning's automatic optimisation calls it) applies the Adam
    # update of every weight matrix inside the epilogue of its weight-gradient GEMM — the gradient never reaches HBM
    # (8 B/parameter less traffic) and the remaining 26 B/parameter move while the tensor cores work on the next tile.
    # Those parameters keep ``grad is None`` for the step (nothing to inspect, clip or accumulate), which is why this is
    # opt-in; results are bit-identical to the unfused step (same gradient bits, same ``adam_one`` arithmetic).
    # Not available with gradient synchronisation between backward and step (data-parallel runs).
    fuse_backward = False

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

    # ------------------------------------------------------------------------------------------------ state / runs
    def _ensure_state(self, flat, p, off):
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
        return st

    def _buffers(self, flat):
        if flat.bf16 is None:
            flat.refresh_bf16()
        if getattr(flat, "adam_m", None) is None:
            flat.adam_m, flat.adam_v = torch.zeros_like(flat.flat), torch.zeros_like(flat.flat)
            flat.adam_hyper = torch.zeros(len(flat.params) + 1, 8, device=flat.device, dtype=torch.float32)
            flat.adam_slot = {}

    def init_all_state(self) -> bool:
        """Create the (zero) state of every parameter now.  torch creates it lazily at a parameter's first gradient;
        step = 0 with zero moments is the same state, and a captured graph must not contain the zero-fills."""
        flat = self._flat()
        if not self._fusable(flat):
            return False
        self._buffers(flat)
        name_of = {id(p): n for n, p in flat.params.items()}
        for group in self.param_groups:
            for p in group["params"]:
                self._ensure_state(flat, p, flat.offsets[name_of[id(p)]])
        return True

    @staticmethod
    def _run_class(name: str) -> str:
        # parameters that can be absent from a step on their own (dropped modality: grad stays None, model.py:158-159)
        # never share a launch with others, so the runs of a captured graph keep a common step count forever
        parts = name.split(".")
        return ".".join(parts[:2]) if parts[0] in ("projectors", "contrastive_heads") else "core"

    def _plan_runs(self, flat, split: bool, only=None, exclude=()):
        """Contiguous [lo, hi) ranges of the flat buffer to update: adjacent parameters with a gradient, the same
        param group and the same step count (and, with ``split``, the same run class) share one launch.  ``only``:
        optional (lo, hi) window of the flat buffer (one gradient bucket)."""
        name_of = {id(p): n for n, p in flat.params.items()}
        flat.ensure_grad()  # gradients assigned by hand (no engine backward yet) are copied into the flat buffer below
        todo = []
        for gi, group in enumerate(self.param_groups):
            for p in group["params"]:
                if p.grad is None:
                    continue
                name = name_of[id(p)]
                off = flat.offsets[name]
                if only is not None and not (only[0] <= off < only[1]):
                    continue
                if any(lo <= off < hi for lo, hi in exclude):
                    continue
                st = self._ensure_state(flat, p, off)
                if p.grad.data_ptr() != flat.grad.data_ptr() + 4 * off:  # foreign gradient tensor: bring it into the flat buffer
                    flat.gview(name).copy_(p.grad)
                todo.append((off, off + _engine._round_up(p.numel(), _engine.ALIGN), int(st["step"].item()), gi,
                             self._run_class(name) if split else "", p))
        todo.sort(key=lambda t: t[0])
        runs = []
        for lo, hi, k, gi, cls_, p in todo:
            if runs and runs[-1]["hi"] == lo and runs[-1]["k"] == k and runs[-1]["group"] == gi and runs[-1]["cls"] == cls_:
                runs[-1]["hi"] = hi
                runs[-1]["params"].append(p)
            else:
                runs.append({"lo": lo, "hi": hi, "k": k, "group": gi, "cls": cls_, "params": [p]})
        return runs

    def _launch(self, flat, run, k, stream, device_hyper: bool, max_blocks: int = 0):
        lib = _lib.load()
        group = self.param_groups[run["group"]]
        lr = float(group["lr"])
        beta1, beta2 = (float(b) for b in group["betas"])
        eps, wd = float(group["eps"]), float(group["weight_decay"])
        lo, n = run["lo"], run["hi"] - run["lo"]
        ptrs = (ctypes.c_void_p(flat.flat.data_ptr() + 4 * lo), ctypes.c_void_p(flat.grad.data_ptr() + 4 * lo),
                ctypes.c_void_p(flat.adam_m.data_ptr() + 4 * lo), ctypes.c_void_p(flat.adam_v.data_ptr() + 4 * lo),
                ctypes.c_void_p(flat.bf16.data_ptr() + 2 * lo))
        if device_hyper:
            slot = flat.adam_slot.setdefault(lo, len(flat.adam_slot))
            check(lib.tribe_adam_step_dev(*ptrs, n, ctypes.c_void_p(flat.adam_hyper[slot].data_ptr()), max_blocks, stream), "tribe_adam_step_dev")
        else:
            check(lib.tribe_adam_step(*ptrs, n, lr, beta1, beta2, eps, wd, k, max_blocks, stream), "tribe_adam_step")

    # ------------------------------------------------------------------------------------------------ optimizer-in-backward
    def arm_fused_backward(self) -> bool:
        """Declare that the backward pass(es) about to run complete ONE optimizer step that ``step()`` will finish:
        the engine's last backward pass of the step then updates the GEMM weights in its wgrad epilogues.  Returns
        whether fusion is active (False: everything goes through ``step()`` as usual)."""
        model = getattr(self, "_tribe_model", None)
        eng = getattr(model, "_engine", None) if model is not None else None
        if eng is None:
            return False
        eng.fused_opt = None
        if not self.fuse_backward or self._tribe_sharded is not None or eng.comm is not None:
            return False
        flat = self._flat()
        if not self._fusable(flat):
            return False
        self._buffers(flat)
        if torch.cuda.is_current_stream_capturing() and getattr(self, "_graph_runs", None) is None:
            raise _lib.TribeError("TribeAdam.arm_fused_backward() inside a CUDA graph capture needs graph_begin()")
        if getattr(self, "_group_of", None) is None or len(self._group_of) != sum(len(g["params"]) for g in self.param_groups):
            self._group_of = {id(p): gi for gi, g in enumerate(self.param_groups) for p in g["params"]}
        eng.fused_opt = self
        return True

    def disarm_fused_backward(self) -> None:
        model = getattr(self, "_tribe_model", None)
        eng = getattr(model, "_engine", None) if model is not None else None
        if eng is not None and eng.fused_opt is self:
            eng.fused_opt = None

    def fused_wgrad_args(self, names):
        """Called by ``engine.Engine._wgrad`` for the adjacent parameters ``names`` whose gradient GEMM is about to be
        launched: does the host bookkeeping of their optimizer step (step counters, device hyper-parameter block) and
        returns the epilogue's pointer tuple, or None when these parameters cannot take the fused path."""
        flat = self._flat()
        params = [flat.params[n] for n in names]
        groups = {self._group_of.get(id(p)) for p in params}
        if len(groups) != 1 or None in groups:
            return None
        gi = groups.pop()
        lo = flat.offsets[names[0]]
        off = lo
        for n, p in zip(names, params):  # one contiguous range
            if flat.offsets[n] != off:
                return None
            off += _engine._round_up(p.numel(), _engine.ALIGN)
        steps = {int(self._ensure_state(flat, p, flat.offsets[n])["step"].item()) for n, p in zip(names, params)}
        if len(steps) != 1:
            return None
        k = steps.pop()
        slot = flat.adam_slot.setdefault(lo, len(flat.adam_slot))
        hyper = flat.adam_hyper[slot].data_ptr()
        if torch.cuda.is_current_stream_capturing():
            self._graph_runs.append({"lo": lo, "hi": off, "k": k, "group": gi, "cls": "fused", "params": params})
        else:
            for p in params:
                self.state[p]["step"] += 1
            group = self.param_groups[gi]
            beta1, beta2 = (float(b) for b in group["betas"])
            check(_lib.load().tribe_adam_hyper(ctypes.c_void_p(hyper), float(group["lr"]), beta1, beta2, float(group["eps"]),
                                               float(group["weight_decay"]), k + 1, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                  "tribe_adam_hyper")
        return (flat.flat.data_ptr() + 4 * lo, flat.adam_m.data_ptr() + 4 * lo, flat.adam_v.data_ptr() + 4 * lo, flat.bf16.data_ptr() + 2 * lo,
                hyper, False)

    # ------------------------------------------------------------------------------------------------ CUDA-graph protocol
    def graph_begin(self) -> None:
        """Called (by graphed.GraphedTrainStep) right before a train step is captured: ``step()`` then records its
        launches with device-resident hyper-parameters and leaves the host-side step counters untouched (a capture
        executes nothing)."""
        self._graph_runs = []

    def graph_end(self) -> list:
        runs, self._graph_runs = self._graph_runs, None
        return runs

    def prepare_replay(self, runs) -> None:
        """Host bookkeeping of one optimizer step whose kernels are about to be replayed: advance the step counters and
        refresh every run's device hyper-parameter block with the scheduler's current lr / betas (one launch)."""
        flat = self._flat()
        lib = _lib.load()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        n = len(runs)
        slots, ks = (ctypes.c_int32 * n)(), (ctypes.c_int64 * n)()
        lr, b1, b2, eps, wd = ((ctypes.c_double * n)() for _ in range(5))
        for i, run in enumerate(runs):
            steps = {int(self.state[p]["step"].item()) for p in run["params"]}
            if len(steps) != 1:
                raise _lib.TribeError("captured Adam run lost its common step count; re-capture the train step")
            for p in run["params"]:
                self.state[p]["step"] += 1
            group = self.param_groups[run["group"]]
            slots[i], ks[i] = flat.adam_slot[run["lo"]], steps.pop() + 1
            lr[i], eps[i], wd[i] = float(group["lr"]), float(group["eps"]), float(group["weight_decay"])
            b1[i], b2[i] = (float(b) for b in group["betas"])
        if n:
            check(lib.tribe_adam_hyper_batch(ctypes.c_void_p(flat.adam_hyper.data_ptr()), slots, ks, lr, b1, b2, eps, wd, n, stream),
                  "tribe_adam_hyper_batch")
        self._opt_called = True  # what LR schedulers look at to warn about scheduler.step() before optimizer.step()
        flat.opt_steps += 1
        flat._sig = (sum(p._version for p in flat.params.values()), flat.opt_steps)

    # ------------------------------------------------------------------------------------------------ step
    def _apply_runs(self, flat, runs, capturing, stream, max_blocks):
        for run in runs:
            if capturing:
                self._launch(flat, run, 0, stream, device_hyper=True, max_blocks=max_blocks)
                self._graph_runs.append(run)
            else:
                for p in run["params"]:
                    self.state[p]["step"] += 1
                self._launch(flat, run, run["k"] + 1, stream, device_hyper=False, max_blocks=max_blocks)

    @torch.no_grad()
    def step_bucket(self, lo_hi, max_blocks: int = 0) -> None:
        """The optimizer step of the parameters inside one gradient bucket [lo, hi) of the flat buffer, on the CURRENT
        stream (parallel.StepOverlap calls this behind the backward pass); the following ``step()`` skips them.
        ``max_blocks`` bounds the grid so the kernel shares the SMs with the backward GEMMs instead of displacing them."""
        flat = self._flat()
        if not self._fusable(flat):
            return  # step() will do everything
        self._buffers(flat)
        capturing = torch.cuda.is_current_stream_capturing()
        if capturing and getattr(self, "_graph_runs", None) is None:
            raise _lib.TribeError("TribeAdam.step_bucket() inside a CUDA graph capture needs graph_begin()")
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        self._apply_runs(flat, self._plan_runs(flat, split=capturing, only=lo_hi), capturing, stream, max_blocks)
        self._early = (getattr(self, "_early", None) or []) + [tuple(lo_hi)]

    @torch.no_grad()
    def step_bucket_sharded(self, idx: int) -> None:
        """Data-parallel step of gradient bucket ``idx`` through ``parallel.ShardedStep`` on the CURRENT stream: the
        fused (in-switch gradient reduction -> Adam -> shadow multicast) kernel on the part of every run this rank owns.
        Host bookkeeping (step counters, hyper-parameter slots) covers the whole bucket on every rank."""
        sh = self._tribe_sharded
        flat = self._flat()
        if not self._fusable(flat):
            raise _lib.TribeError("ShardedStep needs the fused TribeAdam path (no amsgrad / maximize / foreign parameters)")
        self._buffers(flat)
        capturing = torch.cuda.is_current_stream_capturing()
        if capturing and getattr(self, "_graph_runs", None) is None:
            raise _lib.TribeError("TribeAdam.step_bucket_sharded() inside a CUDA graph capture needs graph_begin()")
        lib = _lib.load()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        lo_hi = flat.bucket_ranges[idx]
        own_lo, own_hi = sh.owned[idx][sh.rank]
        for run in self._plan_runs(flat, split=capturing, only=lo_hi):
            slot = flat.adam_slot.setdefault(run["lo"], len(flat.adam_slot))
            hyper = flat.adam_hyper[slot].data_ptr()
            if capturing:
                self._graph_runs.append(run)  # prepare_replay() advances the counters and refreshes the slot
            else:
                for p in run["params"]:
                    self.state[p]["step"] += 1
                group = self.param_groups[run["group"]]
                beta1, beta2 = (float(b) for b in group["betas"])
                check(lib.tribe_adam_hyper(ctypes.c_void_p(hyper), float(group["lr"]), beta1, beta2, float(group["eps"]),
                                           float(group["weight_decay"]), run["k"] + 1, stream), "tribe_adam_hyper")
            a, b = max(run["lo"], own_lo), min(run["hi"], own_hi)
            if b > a:
                for x, y, bcast in sh.pieces(a, b):
                    sh.launch(idx, x, y, bcast, hyper)
        self._early = (getattr(self, "_early", None) or []) + [tuple(lo_hi)]

    def state_dict(self):
        sh = getattr(self, "_tribe_sharded", None)
        if sh is not None:
            sh.gather_optimizer_state()  # collective: Adam moments live on their owner rank only
        return super().state_dict()

    @torch.no_grad()
    def step(self, closure=None):
        flat = self._flat()
        if not self._fusable(flat):
            return super().step(closure)
        loss = None
        if closure is not None:
            # Lightning's automatic optimisation hands training_step + zero_grad + backward over as the closure: this
            # backward completes exactly this step, so it may carry the update of the GEMM weights
            self.arm_fused_backward()
            with torch.enable_grad():
                loss = closure()
        self.disarm_fused_backward()
        self._buffers(flat)
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        capturing = torch.cuda.is_current_stream_capturing()
        if capturing and getattr(self, "_graph_runs", None) is None:
            raise _lib.TribeError("TribeAdam.step() inside a CUDA graph capture needs graph_begin() (see graphed.GraphedTrainStep)")
        early = getattr(self, "_early", None) or ()  # buckets already stepped behind the backward (StepOverlap)
        self._apply_runs(flat, self._plan_runs(flat, split=capturing, exclude=early), capturing, stream, 0)
        self._early = None
        # the shadow weights are already current for the state the post-step hook is about to announce
        flat._sig = (sum(p._version for p in flat.params.values()), flat.opt_steps + 1)
        return loss
