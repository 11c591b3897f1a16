"""Whole-train-step CUDA graphs.

One TRIBE train step is ~245 kernel launches whose arguments (workspace pointers, tensor maps, shapes) do not change
from step to step; enqueueing them from Python costs ~26 ms of host time against ~30 ms of GPU time.  ``GraphedTrainStep``
captures the complete Lightning-style step (``zero_grad -> training_step -> backward -> [gradient all-reduce] ->
Adam``) once per *variant* and replays it afterwards with one ``cudaGraphLaunch``.

What makes a variant (the graph key):
  * the modality-dropout masks of the step — drawn on the HOST with the reference's exact RNG consumption
    (``model.py:134-141``: ``torch.rand(1)`` per modality, NumPy when all are selected) *before* the graph is chosen, so
    masks and generator state stay bit-identical to the eager path; a dropped modality removes its projector GEMMs from
    the graph and keeps that projector out of the optimizer step (``grad is None`` semantics, ``model.py:158-159``);
  * the device addresses / shapes of the batch tensors (``segment.DevicePrefetcher`` cycles two persistent slots).
The first time a variant is seen it runs eagerly (this also warms every lazily allocated buffer), the second time it is
captured, from then on replayed.  Per-step host values that live inside kernels — Adam's lr / betas / bias corrections,
which OneCycleLR changes every batch — are read from device memory (``tribe_adam_step_dev``) and refreshed by one tiny
launch per optimizer run before each replay.  Host-side bookkeeping a replay skips (optimizer step counters, ``p.grad``
presence, the LR scheduler, the launch counter, the subject-range flag) is re-applied explicitly.
"""
from __future__ import annotations

import torch

from . import _lib
from .optim import TribeAdam


class GraphedTrainStep:
    MAX_VARIANTS = 128

    def __init__(self, trainer):
        self.trainer = trainer
        self.cache: dict = {}
        self.captures = 0
        self.replays = 0

    # ------------------------------------------------------------------------------------------------ eligibility
    def _supported(self, batch) -> bool:
        tr = self.trainer
        model = getattr(tr.module, "model", None)
        if model is None or not hasattr(model, "_draw_dropout") or not isinstance(tr.optimizer, TribeAdam):
            return False
        if tr.grad_sync is not None and not tr.graph_collectives:
            from .parallel import ShardedStep, world

            # NCCL all-reduces inside the step are captured only when asked to; the NVLink step tail (our kernels, no
            # NCCL call) always is
            if world()[1] > 1 and not isinstance(tr.grad_sync, ShardedStep):
                return False
        return all(torch.is_tensor(v) and v.is_cuda for v in batch.data.values())

    @staticmethod
    def _key(batch, masks):
        return (tuple(tuple(sorted(str(x) for x in m)) for m in masks),  # a mask is a set (np.random.choice order varies)
                tuple((k, v.data_ptr(), tuple(v.shape), str(v.dtype)) for k, v in sorted(batch.data.items())))

    # ------------------------------------------------------------------------------------------------ the step
    def step(self, batch) -> torch.Tensor:
        tr = self.trainer
        if not self._supported(batch):
            return tr.eager_step(batch)
        model = tr.module.model
        tr.module.train()
        n_draws = 1 + int(bool(getattr(model.config, "contrastive_enabled", False)))
        masks = [model._draw_dropout() for _ in range(n_draws)]  # the step's CPU-RNG draws, in the eager order
        key = self._key(batch, masks)
        entry = self.cache.get(key)
        if entry is None or entry == "seen" and self.captures >= self.MAX_VARIANTS:
            if len(self.cache) >= 8 * self.MAX_VARIANTS:  # batches at ever-new addresses: stop remembering them
                self.cache = {k: v for k, v in self.cache.items() if v != "seen"}
            self.cache[key] = "seen"
            model._preset_dropped = [list(m) for m in masks]
            try:
                return tr.eager_step(batch)
            finally:
                leftover, model._preset_dropped = model._preset_dropped, None
                assert not leftover, "train step consumed fewer dropout draws than were drawn ahead"
        if entry == "seen":
            entry = self.cache[key] = self._capture(batch, masks)
        model.last_dropped = list(masks[-1])  # what the eager path leaves behind after its last aggregate_features call
        return self._replay(entry)

    def warm(self, batches) -> int:
        """Capture every dropout variant of the given device-resident batches now (a capture executes nothing and
        draws nothing: parameters, optimizer state and RNG streams are untouched).  At least one eager step must have
        run before so that every lazily allocated buffer exists.  Returns the number of graphs captured."""
        import itertools

        model = self.trainer.module.model
        mods = list(model.feature_dims.keys())
        p = float(model.config.modality_dropout)
        variants = [[]]
        if p > 0.0:  # every strict subset can be dropped (all-selected falls back to n-1 of them, model.py:138-141)
            variants = [list(c) for r in range(len(mods)) for c in itertools.combinations(mods, r)]
        n_draws = 1 + int(bool(getattr(model.config, "contrastive_enabled", False)))
        n = 0
        was_training = self.trainer.module.training
        self.trainer.module.train()
        for batch in batches:
            if not self._supported(batch):
                continue
            for masks in itertools.product(variants, repeat=n_draws):
                key = self._key(batch, masks)
                if isinstance(self.cache.get(key), dict) or self.captures >= self.MAX_VARIANTS:
                    continue
                self.cache[key] = self._capture(batch, [list(m) for m in masks])
                n += 1
        self.trainer.module.train(was_training)
        return n

    def _capture(self, batch, masks):
        tr = self.trainer
        model, opt = tr.module.model, tr.optimizer
        opt.init_all_state()
        model.flush_subject_check()
        from . import ops

        ops.drain_stale_error("before a CUDA-graph capture")  # a stale error would abort the capture at its first checked call
        model._preset_dropped = [list(m) for m in masks]
        graph = torch.cuda.CUDAGraph()
        lib = _lib.load()
        launches0 = int(lib.tribe_launch_count())
        opt.graph_begin()
        try:
            # thread_local: helper threads (NCCL watchdog, pin-memory, samplers) may keep calling the CUDA runtime
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                loss = tr.run_step_body(batch)
        finally:
            runs = opt.graph_end()
            leftover, model._preset_dropped = model._preset_dropped, None
        assert not leftover, "train step consumed fewer dropout draws than were drawn ahead"
        self.captures += 1
        params = [p for group in opt.param_groups for p in group["params"]]
        return {"graph": graph, "loss": loss, "runs": runs, "grads": [(p, p.grad) for p in params],
                "launches": int(lib.tribe_launch_count()) - launches0, "batch": batch}

    def _replay(self, entry) -> torch.Tensor:
        tr = self.trainer
        tr.optimizer.prepare_replay(entry["runs"])
        entry["graph"].replay()
        if hasattr(tr.grad_sync, "masters_stale"):
            tr.grad_sync.masters_stale = tr.grad_sync.state_stale = True
        for p, g in entry["grads"]:
            p.grad = g
        _lib.REPLAYED_LAUNCHES += entry["launches"]
        if tr.scheduler is not None:
            tr.scheduler.step()
        tr.global_step += 1
        self.replays += 1
        flags = tr.module.model.__dict__.get("_subject_flags")
        if flags is not None and int(flags[1].item()) != 0:
            tr.module.model.flush_subject_check()
        return entry["loss"].clone()  # the graph's static output buffer is overwritten by the next replay
