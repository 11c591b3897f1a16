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


def sharded_pearson(preds_local: torch.Tensor, trues_local: torch.Tensor, group=None) -> torch.Tensor:
    """Per-parcel Pearson r over ALL ranks' rows, parcels sharded across ranks.  Inputs are this rank's (n_local, O)
    row-major matrices or (n_windows_local, O, T) prediction tensors, fp32 CUDA -> (O,) fp32 on every rank."""
    from . import ops

    O = preds_local.shape[1]
    p_shard, _ = exchange_parcel_shards(preds_local, group)
    t_shard, _ = exchange_parcel_shards(trues_local, group)
    stats = torch.zeros(1, 6, p_shard.shape[1], device=p_shard.device, dtype=torch.float64)
    ops.pearson_stats(p_shard.contiguous(), t_shard.contiguous(), stats, layout="no" if p_shard.dim() == 2 else "bdt")
    r, _ = ops.pearson_finalize(stats[0])
    return gather_parcels(r, O, group)


def allreduced_pearson(preds_local: torch.Tensor, trues_local: torch.Tensor, group=None) -> torch.Tensor:
    """Alternative without re-laying the data: local statistics over all parcels + one all-reduce of 6 x O fp64."""
    from . import ops

    stats = torch.zeros(1, 6, preds_local.shape[1], device=preds_local.device, dtype=torch.float64)
    ops.pearson_stats(preds_local.contiguous(), trues_local.contiguous(), stats, layout="no" if preds_local.dim() == 2 else "bdt")
    _, ws = world()
    if ws > 1:
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
