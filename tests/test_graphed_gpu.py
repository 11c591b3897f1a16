"""Whole-step CUDA graphs (algonauts2025_b200/graphed.py) must be indistinguishable from the eager Lightning-style step:
same modality-dropout draws (CPU generator position), same losses, same parameters / Adam state after N steps with
OneCycleLR changing lr and beta1 every batch, same ``grad is None`` pattern for dropped modalities (model.py:158-159)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import algonauts2025_b200  # noqa: E402,F401
from algonauts2025_b200.model import FmriEncoder, FmriEncoderConfig  # noqa: E402
from algonauts2025_b200.pl_module import BrainModule  # noqa: E402
from algonauts2025_b200.segment import SegmentData, synthetic_batch  # noqa: E402
from algonauts2025_b200.trainer import MiniTrainer, default_optimizer  # noqa: E402

SMALL_DIMS = {"text": (2, 96), "audio": (2, 40), "video": (1, 72)}
SMALL = dict(hidden=384, depth=2, heads=6)


def _run(use_graphs: bool, contrastive: bool, steps: int, p_drop: float = 0.4, slots: int = 2, overlap: bool = False):
    torch.manual_seed(21)
    np.random.seed(21)
    cfg = FmriEncoderConfig(n_subjects=3, modality_dropout=p_drop, contrastive_enabled=contrastive)
    model = FmriEncoder(SMALL_DIMS, 200, 25, cfg, **SMALL)
    module = BrainModule(model=model, loss=torch.nn.MSELoss(), optim_config=None, metrics={}, max_epochs=1)
    opt, sched = default_optimizer(model.parameters(), total_steps=steps + 4, lr=3e-3, model=model)
    trainer = MiniTrainer(module, opt, sched, use_graphs=use_graphs, overlap_optimizer=overlap)
    spec = tuple((k, v[0], v[1]) for k, v in SMALL_DIMS.items())
    host = [synthetic_batch(batch_size=3, t=74, t_out=25, n_outputs=200, n_subjects=3, seed=70 + i, dims=spec) for i in range(2)]
    dev = [SegmentData(data={k: v.cuda() for k, v in b.data.items()}, segments=b.segments) for b in host]  # two fixed device slots
    losses, masks, none_grads = [], [], []
    for i in range(steps):
        losses.append(trainer.train_step(dev[i % slots]).clone())
        masks.append(tuple(model.last_dropped))
        none_grads.append(tuple(n for n, p in model.named_parameters() if p.grad is None))
    torch.cuda.synchronize()
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    steps_of = {n: int(opt.state[p]["step"]) for n, p in model.named_parameters() if p in opt.state and len(opt.state[p])}
    return dict(losses=torch.stack(losses).cpu(), masks=masks, none_grads=none_grads, state=state, steps_of=steps_of,
                next_rand=torch.rand(1).item(), lr=opt.param_groups[0]["lr"], trainer=trainer)


@pytest.mark.parametrize("contrastive", [False, True])
def test_graphed_steps_equal_eager_steps(contrastive):
    # the contrastive step draws two masks (forward + brain latents): 49 variants per batch slot, so use one slot
    steps, kw = (28, {}) if not contrastive else (40, dict(p_drop=0.25, slots=1))
    eager = _run(False, contrastive, steps, **kw)
    graphed = _run(True, contrastive, steps, **kw)
    g = graphed["trainer"]._graphed
    assert g.captures >= 2 and g.replays >= steps // 4, (g.captures, g.replays)  # the graphs were really used
    assert eager["masks"] == graphed["masks"]              # same CPU-RNG draws
    assert eager["next_rand"] == graphed["next_rand"]      # ... and the generator ends at the same position
    assert eager["none_grads"] == graphed["none_grads"]    # dropped projectors keep grad None in both
    assert eager["lr"] == graphed["lr"]
    assert any(eager["masks"]) and not all(eager["masks"])
    # parameters that were skipped in some steps carry fewer Adam steps — identically in both modes (a replay of a
    # graph without the text projector must not advance the text projector's step count); state created ahead of a
    # capture (step 0, zero moments) is the only allowed difference
    for n, k in eager["steps_of"].items():
        assert graphed["steps_of"][n] == k, n
    assert all(k == 0 for n, k in graphed["steps_of"].items() if n not in eager["steps_of"])
    # same kernels in the same order on the same data: equal up to the order of fp32 atomics in a few reductions, which
    # lr = 3e-3 Adam steps amplify.  The yardstick is a second EAGER run: graph-vs-eager must not exceed run-to-run noise.
    eager2 = _run(False, contrastive, steps, **kw)
    noise = float((eager2["losses"] - eager["losses"]).abs().max())
    diff = float((graphed["losses"] - eager["losses"]).abs().max())
    assert diff <= max(10.0 * noise, 1e-3 *float(eager["losses"].abs().max())), (diff, noise)
    for k, v in eager["state"].items():
        n_k = float((eager2["state"][k].float() - v.float()).abs().max())
        d_k = float((graphed["state"][k].float() - v.float()).abs().max())
        assert d_k <= max(10.0 * n_k, 1e-3 + 1e-2 *float(v.float().abs().max())), (k, d_k, n_k)


@pytest.mark.parametrize("use_graphs", [False, True])
@pytest.mark.parametrize("contrastive", [False, True])
def test_optimizer_overlapped_with_backward_equals_plain_step(use_graphs, contrastive):
    """parallel.StepOverlap: each layer's Adam launch moves behind the backward (side stream); same arithmetic, same
    step counts, same skipped projectors — also when the whole step is a replayed graph (forked capture streams)."""
    steps, kw = (20, {}) if not contrastive else (30, dict(p_drop=0.25, slots=1))
    plain = _run(False, contrastive, steps, **kw)
    plain2 = _run(False, contrastive, steps, **kw)
    over = _run(use_graphs, contrastive, steps, overlap=True, **kw)
    assert over["trainer"].grad_sync is not None and over["trainer"].grad_sync.opt_stream is not None
    assert plain["masks"] == over["masks"] and plain["none_grads"] == over["none_grads"] and plain["lr"] == over["lr"]
    for n, k in plain["steps_of"].items():
        assert over["steps_of"][n] == k, n
    noise = float((plain2["losses"] - plain["losses"]).abs().max())
    diff = float((over["losses"] - plain["losses"]).abs().max())
    assert diff <= max(10.0 * noise, 1e-3 *float(plain["losses"].abs().max())), (diff, noise)
    for k, v in plain["state"].items():
        n_k = float((plain2["state"][k].float() - v.float()).abs().max())
        d_k = float((over["state"][k].float() - v.float()).abs().max())
        assert d_k <= max(10.0 * n_k, 1e-3 + 1e-2 *float(v.float().abs().max())), (k, d_k, n_k)


def test_graphed_step_falls_back_for_host_batches_and_reports_bad_subjects():
    torch.manual_seed(3)
    model = FmriEncoder(SMALL_DIMS, 200, 25, FmriEncoderConfig(n_subjects=3), **SMALL)
    module = BrainModule(model=model, loss=torch.nn.MSELoss(), optim_config=None, metrics={}, max_epochs=1)
    opt, sched = default_optimizer(model.parameters(), total_steps=50, model=model)
    trainer = MiniTrainer(module, opt, sched, use_graphs=True)
    spec = tuple((k, v[0], v[1]) for k, v in SMALL_DIMS.items())
    host = synthetic_batch(batch_size=3, t=74, t_out=25, n_outputs=200, n_subjects=3, seed=1, dims=spec)
    assert torch.isfinite(trainer.train_step(host))  # CPU batch: eager path (inputs are moved inside forward)
    assert trainer._graphed.captures == 0
    dev = SegmentData(data={k: v.cuda() for k, v in host.data.items()}, segments=host.segments)
    for _ in range(4):
        trainer.train_step(dev)
    assert trainer._graphed.replays >= 2
    dev.data["subject_id"].fill_(7)  # out of range: the replayed graph still checks on the device
    with pytest.raises(AssertionError):
        for _ in range(3):
            trainer.train_step(dev)
        model.flush_subject_check()
