"""Data-parallel training check on real NVLink / NCCL (run under torchrun on >= 2 GPUs; collected by
tests/test_multigpu_gpu.py).  Every rank trains on its own windows.  Paths compared:

  allreduce   parallel.GradAllReduce: bucketed NCCL all-reduce (mean) + replicated fused Adam
  staged      parallel.ShardedStep: copy-engine pushes of every finished bucket's slices to their owners during the backward,
              then our fused kernel (sum of the staged copies -> rank-sharded Adam -> multimem.st of the bf16 shadow),
              eager and replayed as whole-step CUDA graphs
  nvls        the same kernel reducing through multimem.ld_reduce (in-switch sum)
  p2p         the same kernel's peer-pointer variant (no multicast)

After a few Adam steps (a) all ranks hold bit-identical weights (fp32 masters after the lazy gather AND the bf16 shadow
the kernels read), (b) every path agrees with the all-reduce path to run-to-run noise (the small gradients are
accumulated with atomics, so two runs of the SAME path differ in the last bits too), (c) the all-reduce path matches a single-process run
on the concatenated batch up to bf16 noise, (d) with modality dropout the skipped projectors stay skipped on every
rank, (e) Adam moments gathered from their owners equal the replicated ones, (f) no cross-rank wait timed out."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import algonauts2025_b200  # noqa: E402
from algonauts2025_b200 import parallel  # noqa: E402
from algonauts2025_b200.model import FmriEncoder, FmriEncoderConfig  # noqa: E402
from algonauts2025_b200.pl_module import BrainModule  # noqa: E402
from algonauts2025_b200.segment import SegmentData, synthetic_batch  # noqa: E402
from algonauts2025_b200.trainer import MiniTrainer, default_optimizer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
DIMS = {"text": (2, 96), "audio": (2, 40), "video": (1, 72)}
SMALL = dict(hidden=384, depth=2, heads=6)
SPEC = tuple((k, v[0], v[1]) for k, v in DIMS.items())


def batch_of(seed, b=4):
    h = synthetic_batch(batch_size=b, t=74, t_out=25, n_outputs=200, n_subjects=3, seed=seed, dims=SPEC)
    return SegmentData(data={k: v.cuda() for k, v in h.data.items()}, segments=h.segments)


def cat(batches):
    return SegmentData(data={k: torch.cat([b.data[k] for b in batches]) for k in batches[0].data}, segments=sum((b.segments for b in batches), []))


def train(p_drop, path, graphs, batches, steps=6, contrastive=False):
    torch.manual_seed(7), np.random.seed(7)
    cfg = FmriEncoderConfig(n_subjects=3, modality_dropout=p_drop, contrastive_enabled=contrastive)
    model = FmriEncoder(DIMS, 200, 25, cfg, **SMALL)
    model.independent_replica = path == "single"  # the single-process reference run deliberately has no gradient exchange
    module = BrainModule(model=model, loss=torch.nn.MSELoss(), optim_config=None, metrics={}, max_epochs=1)
    opt, sched = default_optimizer(model.parameters(), total_steps=steps + 2, lr=1e-3, model=model)
    if path == "single":
        gs = None
    elif path == "allreduce":
        gs = parallel.GradAllReduce(model)
    else:
        gs = parallel.ShardedStep(model, opt, mode=path)
        if gs.mode != path and rank == 0:
            print(f"note: no multicast mapping on this box — '{path}' runs as '{gs.mode}'", flush=True)
    tr = MiniTrainer(module, opt, sched, grad_sync=gs, use_graphs=graphs, graph_collectives=True)
    losses = [tr.train_step(batches[i % len(batches)]).clone() for i in range(steps)]
    torch.cuda.synchronize()
    if hasattr(gs, "check"):
        gs.check()
    shadow = model._engine.flat.bf16.clone()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}            # lazy gather of the fp32 masters
    osd = opt.state_dict()["state"]                                                # ... and of the Adam moments
    moments = torch.cat([osd[i]["exp_avg"].flatten() for i in sorted(osd)] + [osd[i]["exp_avg_sq"].flatten() for i in sorted(osd)])
    return sd, torch.stack(losses), tr, shadow, moments.clone()


def same_on_all_ranks(t, what):
    others = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(others, t.contiguous())
    assert all(torch.equal(o, others[0]) for o in others), f"ranks diverged: {what}"


mine = [batch_of(100 + 10 * rank + i) for i in range(2)]
everyone = [cat([batch_of(100 + 10 * r + i) for r in range(world)]) for i in range(2)]
for p_drop, contrastive in ((0.0, False), (0.5, False), (0.5, True)):
    # the contrastive step draws two dropout masks (49 variants per batch slot): one slot and 24 steps, so that variants
    # repeat and the graph runs really replay (a variant is captured at its second sighting)
    kw = dict(contrastive=contrastive, steps=24) if contrastive else dict(contrastive=contrastive)
    data = mine[:1] if contrastive else mine
    base_sd, base_losses, _, base_shadow, base_mom = train(p_drop, "allreduce", False, data, **kw)
    flat = torch.cat([v.flatten().float() for v in base_sd.values()])
    same_on_all_ranks(flat, "allreduce masters")
    # noise floor: the SAME path run twice (atomically accumulated small gradients + Adam's normalisation of tiny gradients)
    again_sd, again_losses, _, _, again_mom = train(p_drop, "allreduce", False, data, **kw)

    def deviation(a_sd, b_sd):
        return max(float((a_sd[k].float() - b_sd[k].float()).abs().max()) / (1e-3 + float(b_sd[k].float().abs().max())) for k in a_sd)

    def deviation_l2(a_sd, b_sd):
        return max(float((a_sd[k].float() - b_sd[k].float()).norm()) / (1e-12 + float(b_sd[k].float().norm())) for k in a_sd)

    floor = deviation(again_sd, base_sd)
    floor_l2 = deviation_l2(again_sd, base_sd)
    floor_mom = float((again_mom - base_mom).abs().max()) / float(base_mom.abs().max())
    if p_drop == 0.0:
        ref_sd, _, _, _, _ = train(0.0, "single", False, everyone)
        for k in base_sd:
            d = float((base_sd[k].float() - ref_sd[k].float()).abs().max())
            assert d <= 2e-3 + 2e-2 * float(ref_sd[k].float().abs().max()), (k, d)
    for path, graphs in (("staged", False), ("staged", True), ("nvls", True), ("p2p", True), ("allreduce", True)):
        if contrastive and path == "allreduce":
            continue
        sd, losses, tr, shadow, mom = train(p_drop, path, graphs, data, **kw)
        same_on_all_ranks(torch.cat([v.flatten().float() for v in sd.values()]), f"{path} masters")
        same_on_all_ranks(shadow.view(torch.uint8), f"{path} bf16 shadow")
        if graphs:
            assert tr._graphed.replays >= 1, f"{path}: graphs were not replayed"
        # separate runs differ by the run-to-run noise of the atomically accumulated small gradients (norm gains, residual
        # scales, positional embedding): compare against the all-reduce run like test_graphed_gpu compares graph vs eager
        # Adam moves an element whose gradient is ~0 by ~lr per step in a direction the atomics order decides, so the
        # element-wise maximum is bimodal from run to run (1.8e-2 or 3e-4 for the SAME pair of runs): the criterion is the
        # relative L2 distance per tensor against the noise floor; the element-wise figure is only bounded grossly
        worst, worst_l2 = deviation(sd, base_sd), deviation_l2(sd, base_sd)
        assert worst_l2 <= max(10.0 * floor_l2, 5e-3), (path, graphs, worst_l2, floor_l2)
        assert worst <= max(4.0 * floor, 5e-2), (path, graphs, worst, floor)
        same_on_all_ranks(mom, f"{path} gathered Adam moments")
        dev_mom = float((mom - base_mom).abs().max()) / float(base_mom.abs().max())
        assert dev_mom <= 4.0 * floor_mom + 1e-3, (path, "Adam moments", dev_mom, floor_mom)
        assert float((losses - base_losses).abs().max()) <= 4.0 * float((again_losses - base_losses).abs().max()) + 2e-3 * float(base_losses.abs().max())
        dist.barrier()
        if rank == 0:
            print(f"dp train check p_drop={p_drop} contrastive={contrastive} path={path} graphs={graphs}: ranks bit-identical, "
                  f"relative L2 deviation from the all-reduce run {worst_l2:.1e} (two all-reduce runs differ by {floor_l2:.1e}; element-wise max {worst:.1e} / {floor:.1e}), "
                  f"Adam moments {dev_mom:.1e} ({floor_mom:.1e})", flush=True)

# ---- the Lightning route: BrainModule.configure_optimizers installs the gradient path itself in a multi-rank job;
# torch DDP around the module is a harmless no-op (its reducer never sees a gradient); a bare model fails loudly
class _OptimConfig:  # stand-in for modeling_utils' optimizer config (optimizers/base.py:47-48, 84-96)
    def copy(self):
        return self

    def build(self, params, total_steps):
        return torch.optim.Adam(params, lr=1e-3)


torch.manual_seed(7), np.random.seed(7)
model = FmriEncoder(DIMS, 200, 25, FmriEncoderConfig(n_subjects=3, modality_dropout=0.0), **SMALL)
module = BrainModule(model=model, loss=torch.nn.MSELoss(), optim_config=_OptimConfig(), metrics={}, max_epochs=1)
model.train()
try:
    model(mine[0])
    raise AssertionError("a training forward without gradient synchronisation must raise in a multi-rank job")
except algonauts2025_b200.TribeError:
    pass
opt = module.configure_optimizers()
assert module._grad_sync is not None, "configure_optimizers did not install the data-parallel gradient path"
ddp = torch.nn.parallel.DistributedDataParallel(module, device_ids=[local], find_unused_parameters=True)  # what Lightning's strategy builds
module.train()
for i in range(4):  # Lightning's automatic optimisation order
    opt.zero_grad(set_to_none=True)
    loss = module.training_step(mine[i % 2], i)
    loss.backward()
    module.on_before_optimizer_step(opt)
    opt.step()
torch.cuda.synchronize()
if hasattr(module._grad_sync, "check"):
    module._grad_sync.check()
sd = model.state_dict()
same_on_all_ranks(torch.cat([v.flatten().float() for v in sd.values()]), "BrainModule-installed gradient path")
same_on_all_ranks(model._engine.flat.bf16.view(torch.uint8), "BrainModule-installed gradient path, shadow")
del ddp
if rank == 0:
    print(f"BrainModule.configure_optimizers route ({type(module._grad_sync).__name__}) under a DDP wrapper: ranks bit-identical", flush=True)
    print("multi-GPU data-parallel training OK", flush=True)
sys.stdout.flush()
os._exit(0)  # graphs that captured NCCL work keep the communicator busy at teardown
