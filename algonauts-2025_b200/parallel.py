"""Multi-GPU plumbing for the hot path: one process per GPU, ``torch.distributed`` (NCCL over NVLink 5 / NVSwitch).

* ``StepOverlap`` (alias ``GradAllReduce``)
                          data-parallel training (reference: Lightning DDP, algonauts2025/main.py:388-394): each layer's
                          contiguous gradient bucket is all-reduced (mean) as soon as that layer's backward has written
                          it, overlapping the rest of the backward.
* ``parcel_bounds`` / ``exchange_parcel_shards`` / ``sharded_pearson``
                          evaluation sharded by parcel (BASELINE.json config 5): every rank predicts its own windows,
                          one all-to-all re-lays (windows-shard x all parcels) into (all windows x parcel-shard), each
                          rank reduces its parcels with the Pearson kernel, r is all-gathered.
* ``ensemble_weights`` / ``ensemble_average``
                          one ensemble member per GPU (algonauts2025/grids/average_submissions.py:107-125): per-parcel
                          softmax(r / tau) weights, weighted all-reduce of the members' predictions.
Shard-exchange logic is backend-agnostic (tested on CPU with gloo, world_size 2); the reductions are CUDA kernels.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


# ------------------------------------------------------------------------------------------------ data-parallel grads
class StepOverlap:
    """Per-bucket pipeline behind the backward pass.  ``engine.backward`` calls ``bucket_ready(i)`` when bucket ``i`` of
    the flat gradient buffer (0 = head, 1..depth = encoder layers) has received its last contribution of the step; then

    * data-parallel runs (world size > 1): the bucket (453 MB fp32 per layer) is all-reduced (mean) on NCCL's stream
      while the main stream continues with the backward of earlier layers (reference: Lightning DDP,
      algonauts2025/main.py:388-394);
    * with ``optimizer`` (a ``TribeAdam``): the fused Adam + bf16-shadow kernel of that layer runs on a side stream right
      after (its all-reduce, if any) — 30 B/parameter of pure HBM traffic hidden behind the tensor-bound backward GEMMs
      of the remaining layers.  ``optimizer.step()`` afterwards only handles the head bucket.  Nothing on the main
      stream touches a finished layer again within the step, and ``finish_step`` joins the side streams before the
      next forward.  The arithmetic is exactly ``optimizer.step()``'s; only its position in time moves.
    """

    def __init__(self, model, group=None, optimizer=None, adam_blocks: int = 148, gemm_sms_during_comm: int = 0):
        self.model = model
        self.engine = model._engine
        self.engine.comm = self
        self.group = group
        self.optimizer = optimizer
        self.adam_blocks = adam_blocks
        # > 0: once the first gradient bucket of a step is on the wire, the remaining backward GEMMs run on this many
        # SMs so that the collective's CTAs and the persistent GEMM grid are co-resident (reset at finish_step)
        self.gemm_sms_during_comm = gemm_sms_during_comm
        self._limited = False
        self.opt_stream = None
        self.works = []
        self.passes, self.head_passes = 1, 0
        self.counts = {}
        self.early = []

    def begin_step(self, backward_passes: int | None = None, head_passes: int | None = None):
        """``backward_passes``: encoder backward passes feeding this step's gradients (2 with the contrastive branch,
        pl_module.py:59-77); ``head_passes``: extra writers of the head bucket (one per contrastive head).  A bucket is
        complete after its last contribution.  Defaults are derived from the model's config."""
        cfg = getattr(self.model, "config", None)
        contrastive = bool(getattr(cfg, "contrastive_enabled", False))
        if backward_passes is None:
            backward_passes = 2 if contrastive else 1
        if head_passes is None:
            head_passes = len(getattr(self.model, "contrastive_heads", ())) if contrastive else 0
        self.works, self.counts, self.passes, self.head_passes, self.early = [], {}, backward_passes, head_passes, []

    def bucket_ready(self, idx: int):
        _, ws = world()
        early_adam = self.optimizer is not None and idx >= 1
        if ws == 1 and not early_adam:
            return
        self.counts[idx] = self.counts.get(idx, 0) + 1
        if self.counts[idx] < self.passes + (self.head_passes if idx == 0 else 0):
            return
        start, end = self.engine.flat.bucket_ranges[idx]
        work = None
        if ws > 1:
            g = self.engine.flat.grad[start:end]
            work = dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            self.works.append(work)
            if self.gemm_sms_during_comm > 0 and not self._limited:
                from . import ops

                ops.gemm_set_sm_limit(self.gemm_sms_during_comm)
                self._limited = True
        if early_adam:
            main = torch.cuda.current_stream()
            if self.opt_stream is None:
                self.opt_stream = torch.cuda.Stream()
            self.opt_stream.wait_stream(main)  # the bucket's gradients are final on the main stream
            with torch.cuda.stream(self.opt_stream):
                if work is not None:
                    work.wait()
                self.optimizer.step_bucket((start, end), max_blocks=self.adam_blocks)
            self.early.append((start, end))

    def finish_step(self):
        if self._limited:
            from . import ops

            ops.gemm_set_sm_limit(0)
            self._limited = False
        for w in self.works:
            w.wait()
        self.works = []
        if self.early:
            torch.cuda.current_stream().wait_stream(self.opt_stream)


GradAllReduce = StepOverlap  # the data-parallel use of the same pipeline


# ------------------------------------------------------------------------------------------------ NVLink step tail
BF16_ONLY = ("projectors.", "contrastive_heads.", "encoder.layers.")  # ... whose ".weight" matrices the kernels read as bf16


def _fp32_consumed(name: str) -> bool:
    """True for parameters the kernels read as fp32 (biases, norm gains, residual scales, positional / subject
    embeddings, readout bias): the owner broadcasts their fp32 value; weight matrices travel as bf16 shadow only."""
    if name == "predictor.weights":
        return False
    return not (name.startswith(BF16_ONLY) and name.endswith(".weight"))


class ShardedStep:
    """Data-parallel step tail on NVLink 5 / NVSwitch with OUR kernel instead of NCCL all-reduce + replicated Adam
    (reference: Lightning DDP + ``torch.optim.Adam`` on every rank, algonauts2025/main.py:388-394).

    Every bucket of the flat layout (0 = head, 1..depth = encoder layers) is statically cut into ``world`` contiguous
    slices.  The bf16 shadow, the fp32 masters, a gradient staging buffer and a block of flag words live in symmetric
    memory (``torch.distributed._symmetric_memory``: same allocation on every rank, peer-mapped, NVLS multicast mapping
    when the fabric has one).  ``mode``:

    ``"staged"`` (default)  when ``engine.backward`` reports a bucket complete, this rank's COPY ENGINES push the slice each
        peer owns into that peer's staging buffer (``tribe_memcpy_async``; no SM, no barrier — the slot is free since the
        previous step's closing barrier) while the backward pass of the earlier layers keeps every SM.  After the
        backward: one cross-rank barrier, then ``tribe_sharded_adam_step`` on the owned slices — sum of the ``world``
        staged copies in rank order from local HBM -> Adam -> ``multimem.st`` of the bf16 shadow (and of the fp32 value of
        parameters the kernels read as fp32) into every rank's copy — and a closing barrier.
    ``"nvls"``   nothing moves during the backward; the same kernel reduces through ``multimem.ld_reduce`` (in-switch).
    ``"p2p"``    the same kernel reads the peers' gradient buffers through peer-mapped pointers and stores to every peer
        (no multicast needed).

    No NCCL call is part of the step, so whole steps are captured in CUDA graphs for N > 1 as well.  Adam moments exist
    only on the owner of a slice, and the fp32 master of weight matrices is current only there; ``gather_masters()`` /
    ``gather_optimizer_state()`` (collective, lazy — ``state_dict()`` calls them) complete them.  The mean is the sum
    times 1/N; every rank receives the SAME bits for every parameter by construction."""

    SLOT_REDUCE, SLOT_DONE = 1, 2

    def __init__(self, model, optimizer, group=None, max_blocks: int = 0, mode: str | None = None, use_multicast: bool | None = None,
                 timeout_s: float = 30.0):
        import ctypes
        import os

        import torch.distributed._symmetric_memory as symm

        from . import _lib

        self.model, self.optimizer = model, optimizer
        self.engine = model._engine
        self.engine._check_flat()
        flat = self.flat = self.engine.flat
        self.rank, self.world = world()
        if self.world < 2 or self.world > _lib.XGPU_MAX_WORLD:
            raise _lib.TribeError(f"ShardedStep needs 2..{_lib.XGPU_MAX_WORLD} ranks, got {self.world}")
        if not hasattr(optimizer, "step_bucket_sharded"):
            raise _lib.TribeError("ShardedStep needs a TribeAdam optimizer (TribeAdam.adopt)")
        self.group = group if group is not None else dist.group.WORLD
        self.max_blocks, self.timeout_s = max_blocks, timeout_s
        self._ctypes = ctypes
        dev = flat.device
        # static ownership: bucket b = [s, e) -> rank r owns [s + r * c, min(e, s + (r + 1) * c)), c a multiple of 8 elements
        self.owned = []
        for s, e in flat.bucket_ranges:
            c = -(-(e - s) // self.world)
            c = (c + 7) // 8 * 8
            self.owned.append([(min(e, s + r * c), min(e, s + (r + 1) * c)) for r in range(self.world)])
        # staging layout at owner r: [source rank][bucket-major concatenation of r's owned slices]
        self.prefix = [[0] * len(self.owned) for _ in range(self.world)]
        totals = []
        for r in range(self.world):
            off = 0
            for bi, owners in enumerate(self.owned):
                self.prefix[r][bi] = off
                off += owners[r][1] - owners[r][0]
            totals.append(off)
        self.cap = (max(totals) + 63) // 64 * 64
        flat.rehome(lambda n, dtype: symm.empty(n, dtype=dtype, device=dev))
        self.flags = symm.empty(_lib.XGPU_SLOTS * _lib.XGPU_MAX_WORLD, dtype=torch.int32, device=dev)
        self.flags.zero_()
        self.stage = symm.empty(self.world * self.cap, dtype=torch.float32, device=dev)
        torch.cuda.synchronize(dev)
        self.handles = {k: symm.rendezvous(t, self.group)
                        for k, t in (("grad", flat.grad), ("bf16", flat.bf16), ("flat", flat.flat), ("flags", self.flags), ("stage", self.stage))}
        self.has_multicast = all(int(self.handles[k].multicast_ptr) != 0 for k in ("grad", "bf16", "flat"))
        mode = mode or os.environ.get("TRIBE_DP_MODE") or "staged"
        if use_multicast is False and mode == "nvls":
            mode = "p2p"
        if mode not in ("staged", "nvls", "p2p"):
            raise _lib.TribeError(f"unknown ShardedStep mode {mode!r}")
        if mode == "nvls" and not self.has_multicast:
            mode = "p2p"
        self.mode = mode
        self.out_multicast = self.has_multicast and mode != "p2p" and use_multicast is not False
        self.err = torch.zeros(1, device=dev, dtype=torch.int32)
        self._peers = {k: [int(x) for x in h.buffer_ptrs] for k, h in self.handles.items()}
        self._flag_peers = _lib.TribeXgpuPeers()
        for r in range(self.world):
            self._flag_peers.ptr[r] = self._peers["flags"][r]
        # maximal flat ranges of parameters the kernels read as fp32
        self.bcast = []
        for n in sorted(flat.offsets, key=flat.offsets.get):
            if _fp32_consumed(n):
                lo = flat.offsets[n]
                hi = lo + (flat.params[n].numel() + 63) // 64 * 64
                if self.bcast and self.bcast[-1][1] == lo:
                    self.bcast[-1][1] = hi
                else:
                    self.bcast.append([lo, hi])
        self.stream = torch.cuda.Stream(dev)
        self.counts, self.passes, self.head_passes = {}, 1, 0
        self.ready = []
        self.masters_stale = self.state_stale = False
        self.engine.comm = self
        flat.sharded = self
        optimizer._tribe_sharded = self
        dist.barrier(group=self.group)

    @property
    def multicast(self) -> bool:
        return self.mode == "nvls" or self.out_multicast

    def describe(self) -> str:
        grads = {"staged": "copy-engine pushes of every finished bucket's slices into the owners' staging buffers during the backward pass",
                 "nvls": "NVLS multimem.ld_reduce (in-switch sum)", "p2p": "peer-pointer loads"}[self.mode]
        out = "multimem.st multicast" if self.out_multicast else "per-peer stores"
        return f"{grads} -> rank-sharded fused Adam -> bf16 shadow {out}, one kernel per owned range after the backward (csrc/xgpu.cu), no NCCL in the step"

    # -------------------------------------------------------------------------------------------- kernel plumbing
    def barrier(self, slot: int) -> None:
        from . import _lib

        st = self._ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.load().tribe_xgpu_barrier(self._ctypes.byref(self._flag_peers), self.rank, self.world, slot,
                                                  self._ctypes.c_void_p(self.err.data_ptr()), float(self.timeout_s), st), "tribe_xgpu_barrier")

    def pieces(self, lo: int, hi: int):
        """[lo, hi) cut at the borders of the fp32-consumed ranges -> (lo, hi, bcast) pieces."""
        cuts = {lo, hi}
        for a, b in self.bcast:
            for x in (a, b):
                if lo < x < hi:
                    cuts.add(x)
        cuts = sorted(cuts)
        out = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            out.append((a, b, any(x <= a and b <= y for x, y in self.bcast)))
        return out

    def push_bucket(self, idx: int) -> None:
        """Copy-engine pushes of bucket ``idx``: the slice rank r owns goes to r's staging area for source ``self.rank``."""
        from . import _lib

        lib, fl = _lib.load(), self.flat
        st = self._ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for k in range(1, self.world):
            r = (self.rank + k) % self.world  # stagger the destinations so that the ranks do not all hit the same peer first
            lo, hi = self.owned[idx][r]
            if hi > lo:
                dst = self._peers["stage"][r] + 4 * (self.rank * self.cap + self.prefix[r][idx])
                _lib.check(lib.tribe_memcpy_async(self._ctypes.c_void_p(dst), self._ctypes.c_void_p(fl.grad.data_ptr() + 4 * lo), 4 * (hi - lo), st),
                           "tribe_memcpy_async")

    def kernel_args(self, idx: int, lo: int, hi: int, bcast: bool, hyper_ptr: int):
        """Arguments of the fused kernel for [lo, hi) (inside this rank's slice of bucket ``idx``)."""
        from . import _lib

        fl = self.flat
        a = _lib.TribeShardedAdam()
        a.param, a.m, a.v = fl.flat.data_ptr() + 4 * lo, fl.adam_m.data_ptr() + 4 * lo, fl.adam_v.data_ptr() + 4 * lo
        a.hyper = hyper_ptr
        if self.mode == "nvls":
            a.grad_mc = int(self.handles["grad"].multicast_ptr) + 4 * lo
        if self.out_multicast:
            a.shadow_mc = int(self.handles["bf16"].multicast_ptr) + 2 * lo
            a.param_mc = int(self.handles["flat"].multicast_ptr) + 4 * lo
        own_lo = self.owned[idx][self.rank][0]
        for r in range(self.world):
            if self.mode == "staged":
                a.grad_peer.ptr[r] = (fl.grad.data_ptr() + 4 * lo if r == self.rank else
                                      self.stage.data_ptr() + 4 * (r * self.cap + self.prefix[self.rank][idx] + (lo - own_lo)))
            else:
                a.grad_peer.ptr[r] = self._peers["grad"][r] + 4 * lo
            a.shadow_peer.ptr[r] = self._peers["bf16"][r] + 2 * lo
            a.param_peer.ptr[r] = self._peers["flat"][r] + 4 * lo
        a.n, a.world, a.rank, a.bcast_master, a.max_blocks = hi - lo, self.world, self.rank, int(bcast), self.max_blocks
        return a

    def launch(self, idx: int, lo: int, hi: int, bcast: bool, hyper_ptr: int) -> None:
        """The fused reduce -> Adam -> multicast kernel on [lo, hi) of the flat layout (a range this rank owns)."""
        from . import _lib

        a = self.kernel_args(idx, lo, hi, bcast, hyper_ptr)
        st = self._ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.load().tribe_sharded_adam_step(self._ctypes.byref(a), st), "tribe_sharded_adam_step")

    # -------------------------------------------------------------------------------------------- step protocol
    def begin_step(self, backward_passes: int | None = None, head_passes: int | None = None):
        cfg = getattr(self.model, "config", None)
        contrastive = bool(getattr(cfg, "contrastive_enabled", False))
        self.passes = backward_passes if backward_passes is not None else (2 if contrastive else 1)
        self.head_passes = head_passes if head_passes is not None else (len(getattr(self.model, "contrastive_heads", ())) if contrastive else 0)
        self.counts, self.ready = {}, []

    def bucket_ready(self, idx: int):
        self.counts[idx] = self.counts.get(idx, 0) + 1
        if self.counts[idx] < self.passes + (self.head_passes if idx == 0 else 0):
            return
        self.ready.append(idx)
        if self.mode == "staged":
            self.stream.wait_stream(torch.cuda.current_stream())  # this rank's contributions to the bucket are final
            with torch.cuda.stream(self.stream):
                self.push_bucket(idx)

    def finish_step(self):
        """After the backward pass: every rank's gradients are in place (barrier), the owned slices are reduced, stepped and
        multicast, and a closing barrier says that all ranks have read every gradient and written every shadow slice —
        the next forward may start and the next backward may overwrite the gradient / staging buffers."""
        main = torch.cuda.current_stream()
        if self.mode == "staged":
            main.wait_stream(self.stream)      # this rank's pushes have landed
        self.barrier(self.SLOT_REDUCE)         # ... and so have everyone else's
        for idx in self.ready:
            self.optimizer.step_bucket_sharded(idx)
        self.barrier(self.SLOT_DONE)
        if self.ready:
            self.masters_stale = self.state_stale = True

    def check(self) -> None:
        """Raise if a cross-rank wait timed out (synchronises the device)."""
        code = int(self.err.item())
        if code:
            from ._lib import TribeError

            raise TribeError(f"cross-GPU barrier timed out at slot {code - 1} on rank {self.rank}: a peer never arrived")

    # -------------------------------------------------------------------------------------------- lazy completion
    def _gather(self, buffers) -> None:
        for owners in self.owned:
            for r, (lo, hi) in enumerate(owners):
                if hi > lo:
                    for buf in buffers:
                        dist.broadcast(buf[lo:hi], src=dist.get_global_rank(self.group, r), group=self.group)

    def gather_masters(self) -> None:
        """Every rank receives the owners' fp32 masters (collective; called by ``state_dict()`` and before a re-cast)."""
        if self.masters_stale:
            self._gather([self.flat.flat])
            self.masters_stale = False

    def gather_optimizer_state(self) -> None:
        """Every rank receives the owners' Adam moments (collective; called by ``TribeAdam.state_dict()``)."""
        if self.state_stale and getattr(self.flat, "adam_m", None) is not None:
            self._gather([self.flat.adam_m, self.flat.adam_v])
            self.state_stale = False


def data_parallel(model, optimizer, group=None, prefer: str = "nvlink", **kw):
    """The gradient-synchronisation object for ``trainer.MiniTrainer(grad_sync=...)``: ``ShardedStep`` (copy-engine gradient
    pushes + our fused reduce / Adam / multicast kernel, rank-sharded optimizer) when symmetric memory can be set up on
    every rank, else ``GradAllReduce`` (NCCL all-reduce + replicated Adam).  The decision is collective."""
    rank, ws = world()
    if ws == 1:
        return None
    ok, sync, why = 0, None, ""
    if prefer == "nvlink" and hasattr(optimizer, "step_bucket_sharded"):
        try:
            sync = ShardedStep(model, optimizer, group=group, **kw)
            ok = 1
        except Exception as e:  # noqa: BLE001 - any failure of the symmetric-memory setup selects the NCCL path
            why = f"{type(e).__name__}: {e}"
    flag = torch.tensor([ok], device="cuda", dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 1:
        return sync
    if sync is not None:  # another rank failed: undo
        model._engine.comm = None
        model._engine.flat.sharded = None
        optimizer._tribe_sharded = None
    if rank == 0 and prefer == "nvlink":
        print(f"[tribe] NVLink step tail unavailable ({why or 'a peer failed'}); using NCCL all-reduce", flush=True)
    return GradAllReduce(model, group=group)


# ------------------------------------------------------------------------------------------------ parcel-sharded eval
def parcel_bounds(n_parcels: int, world_size: int) -> list[tuple[int, int]]:
    """Contiguous, balanced parcel shards (first ``n % G`` shards get one extra parcel)."""
    base, extra = divmod(n_parcels, world_size)
    bounds, lo = [], 0
    for r in range(world_size):
        hi = lo + base + (1 if r < extra else 0)
        bounds.append((lo, hi))
        lo = hi
    return bounds


def exchange_parcel_shards(local: torch.Tensor, group=None) -> tuple[torch.Tensor, tuple[int, int]]:
    """local: (n_local_rows, O) rows — or (n_local_windows, O, T) predictions as the model returns them — owned by this
    rank -> (n_total, O_shard[, T]): every rank's rows/windows for THIS rank's parcel shard, ordered by source rank.
    One all-gather of counts + one all-to-all."""
    rank, ws = world()
    n_local, O = local.shape[0], local.shape[1]
    tail = tuple(local.shape[2:])
    bounds = parcel_bounds(O, ws)
    lo, hi = bounds[rank]
    if ws == 1:
        return local, (lo, hi)
    counts = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(ws)]
    dist.all_gather(counts, torch.tensor([n_local], dtype=torch.int64, device=local.device), group=group)
    counts = [int(c.item()) for c in counts]
    send = [local[:, a:b].contiguous() for a, b in bounds]
    recv = [torch.empty(counts[src], hi - lo, *tail, dtype=local.dtype, device=local.device) for src in range(ws)]
    if dist.get_backend(group) == "nccl":
        dist.all_to_all(recv, send, group=group)
    else:  # gloo (CPU tests) has no all-to-all: the same blocks move as point-to-point sends
        recv[rank].copy_(send[rank])
        reqs = [dist.isend(send[dst], dst, group=group) for dst in range(ws) if dst != rank]
        reqs += [dist.irecv(recv[src], src, group=group) for src in range(ws) if src != rank]
        for req in reqs:
            req.wait()
    return torch.cat(recv, dim=0), (lo, hi)


def gather_parcels(r_shard: torch.Tensor, n_parcels: int, group=None) -> torch.Tensor:
    rank, ws = world()
    if ws == 1:
        return r_shard
    bounds = parcel_bounds(n_parcels, ws)
    width = max(b - a for a, b in bounds)  # shards differ by at most one parcel: pad to a common width
    mine = torch.zeros(width, dtype=r_shard.dtype, device=r_shard.device)
    mine[: r_shard.numel()] = r_shard
    parts = [torch.empty(width, dtype=r_shard.dtype, device=r_shard.device) for _ in bounds]
    dist.all_gather(parts, mine, group=group)
    return torch.cat([p[: b - a] for p, (a, b) in zip(parts, bounds)])


def sharded_pearson(preds_local: torch.Tensor, trues_local: torch.Tensor, group=None, timings: dict | None = None) -> torch.Tensor:
    """Per-parcel Pearson r over ALL ranks' rows, parcels sharded across ranks.  Inputs are this rank's (n_local, O)
    row-major matrices or (n_windows_local, O, T) prediction tensors, fp32 CUDA -> (O,) fp32 on every rank.
    ``timings`` (optional dict) receives CUDA events around the exchange / statistics / gather stages."""
    from . import ops

    def stamp(name):
        if timings is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            timings.setdefault(name, []).append(e)

    O = preds_local.shape[1]
    stamp("exchange")
    p_shard, _ = exchange_parcel_shards(preds_local, group)
    t_shard, _ = exchange_parcel_shards(trues_local, group)
    stamp("exchange")
    stamp("pearson")
    r, _ = ops.pearson_r(p_shard.contiguous(), t_shard.contiguous(), layout="no" if p_shard.dim() == 2 else "bdt")
    stamp("pearson")
    stamp("gather")
    out = gather_parcels(r, O, group)
    stamp("gather")
    return out


def allreduced_pearson(preds_local: torch.Tensor, trues_local: torch.Tensor, group=None) -> torch.Tensor:
    """Alternative without re-laying the data: local statistics over all parcels + one all-reduce of 6 x O fp64."""
    from . import ops

    layout = "no" if preds_local.dim() == 2 else "bdt"
    stats = torch.zeros(1, 6, preds_local.shape[1], device=preds_local.device, dtype=torch.float64)
    shift = torch.zeros(2, preds_local.shape[1], device=preds_local.device, dtype=torch.float32)
    if preds_local.numel():
        ops.pearson_pick_shift(preds_local.contiguous(), trues_local.contiguous(), shift, layout=layout)
        ops.pearson_stats(preds_local.contiguous(), trues_local.contiguous(), stats, layout=layout, shift=shift)
    _, ws = world()
    if ws > 1:
        ops.pearson_recenter(stats, shift, None)  # ranks picked different pivots: merge about pivot 0 in fp64
        dist.all_reduce(stats, group=group)
    r, _ = ops.pearson_finalize(stats[0])
    return r


# ------------------------------------------------------------------------------------------------ ensembles
def ensemble_weights(member_r: torch.Tensor, temperature: float = 0.3, softmax_over: str = "voxels") -> torch.Tensor:
    """member_r (N, O) -> (N, O) weights.  ``"voxels"``: ``softmax(r / tau, dim=1)`` — the reference's arithmetic
    (average_submissions.py:108-109 normalises every member's weights over the voxel axis); ``"members"``: softmax over the
    members per parcel (a convex combination per parcel)."""
    return torch.softmax(member_r / temperature, dim=1 if softmax_over == "voxels" else 0)


def ensemble_average(pred_member: torch.Tensor, r_member: torch.Tensor, temperature: float = 0.3, group=None,
                     softmax_over: str = "voxels") -> torch.Tensor:
    """One member per rank: pred_member (..., O, T) or (N_rows, O) predictions of THIS rank's model on the shared
    evaluation set, r_member (O,) its per-parcel validation Pearson.  Returns the weighted ensemble prediction (identical
    on every rank): [all-gather of r (O floats per rank) when the weights are normalised over members] + one all-reduce
    of the weighted predictions."""
    rank, ws = world()
    if softmax_over == "voxels":
        w = ensemble_weights(r_member[None, :], temperature, "voxels")[0]  # depends on this member's r only
    else:
        rs = [torch.empty_like(r_member) for _ in range(ws)]
        if ws > 1:
            dist.all_gather(rs, r_member.contiguous(), group=group)
        else:
            rs = [r_member]
        w = ensemble_weights(torch.stack(rs), temperature, "members")[rank]  # (O,)
    parcel_dim = 1
    shape = [1] * pred_member.dim()
    shape[parcel_dim] = -1
    out = pred_member * w.view(shape)
    if ws > 1:
        dist.all_reduce(out, group=group)
    return out
