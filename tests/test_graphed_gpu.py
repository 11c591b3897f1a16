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


def _run(use_graphs: bool, contrastive: bool, steps: int, p_drop: float = 0.4, slots: int = 2, overlap: bool = False, lr: float = 3e-3,
         fuse=None):
    torch.manual_seed(21)
    np.random.seed(21)
    cfg = FmriEncoderConfig(n_subjects=3, modality_dropout=p_drop, contrastive_enabled=contrastive)
    model = FmriEncoder(SMALL_DIMS, 200, 25, cfg, **SMALL)
    module = BrainModule(model=model, loss=torch.nn.MSELoss(), optim_config=None, metrics={}, max_epochs=1)
    opt, sched = default_optimizer(model.parameters(), total_steps=steps + 4, lr=lr, model=model)
    trainer = MiniTrainer(module, opt, sched, use_graphs=use_graphs, overlap_optimizer=overlap, fuse_optimizer=fuse)
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
    moments = {n: (opt.state[p]["exp_avg"].detach().clone(), opt.state[p]["exp_avg_sq"].detach().clone())
               for n, p in model.named_parameters() if p in opt.state and len(opt.state[p])}
    shadow = model._engine.flat.bf16.detach().clone()
    return dict(moments=moments, shadow=shadow, losses=torch.stack(losses).cpu(), masks=masks, none_grads=none_grads, state=state, steps_of=steps_of,
                next_rand=torch.rand(1).item(), lr=opt.param_groups[0]["lr"], trainer=trainer)


def _assert_same_bookkeeping(ref, got):
    """Everything discrete must agree exactly: CPU-RNG draws, generator position, grad-None pattern, lr, Adam step counts."""
    assert ref["masks"] == got["masks"]
    assert ref["next_rand"] == got["next_rand"]
    assert ref["none_grads"] == got["none_grads"]
    assert ref["lr"] == got["lr"]
    for n, k in ref["steps_of"].items():
        assert got["steps_of"][n] == k, n
    # state created ahead of a capture (step 0, zero moments) is the only allowed difference
    assert all(k == 0 for n, k in got["steps_of"].items() if n not in ref["steps_of"])


def _assert_same_numbers(ref, ref2, got):
    """Same kernels in the same order on the same data: equal up to the order of fp32 atomics in a few reductions.  The
    yardstick is a second run of the reference mode (run-to-run noise); short horizon + small lr keep Adam from amplifying it."""
    noise = float((ref2["losses"] - ref["losses"]).abs().max())
    diff = float((got["losses"] - ref["losses"]).abs().max())
    assert diff <= max(10.0 * noise, 5e-4 * float(ref["losses"].abs().max())), (diff, noise)
    # parameters are compared in the L2 sense: Adam moves an element with a near-zero gradient by ~lr in a direction that
    # atomics-order noise decides, so single elements may differ by a few lr even between two runs of the same mode
    for k, v in ref["state"].items():
        scale = float(v.float().norm()) + 1e-12
        n_k = float((ref2["state"][k].float() - v.float()).norm()) / scale
        d_k = float((got["state"][k].float() - v.float()).norm()) / scale
        assert d_k <= max(10.0 * n_k, 1e-2), (k, d_k, n_k)


@pytest.mark.parametrize("contrastive", [False, True])
def test_graphed_steps_keep_the_eager_bookkeeping_under_modality_dropout(contrastive):
    # the contrastive step draws two masks (forward + brain latents): 49 variants per batch slot, so use one slot
    steps, kw = (28, {}) if not contrastive else (40, dict(p_drop=0.25, slots=1))
    eager = _run(False, contrastive, steps, **kw)
    graphed = _run(True, contrastive, steps, **kw)
    g = graphed["trainer"]._graphed
    assert g.captures >= 2 and g.replays >= steps // 4, (g.captures, g.replays)  # the graphs were really used
    assert any(eager["masks"]) and not all(eager["masks"])
    _assert_same_bookkeeping(eager, graphed)
    # 28-40 lr = 3e-3 Adam steps amplify atomics-order noise chaotically: only a gross check here, the tight one is below
    assert float((graphed["losses"] - eager["losses"]).abs().max()) <= 0.05 * float(eager["losses"].abs().max())


@pytest.mark.parametrize("contrastive", [False, True])
def test_graphed_steps_equal_eager_steps(contrastive):
    kw = dict(p_drop=0.0, slots=2, lr=5e-4)
    eager, eager2 = _run(False, contrastive, 8, **kw), _run(False, contrastive, 8, **kw)
    graphed = _run(True, contrastive, 8, **kw)
    assert graphed["trainer"]._graphed.replays >= 4
    _assert_same_bookkeeping(eager, graphed)
    _assert_same_numbers(eager, eager2, graphed)


@pytest.mark.parametrize("use_graphs", [False, True])
@pytest.mark.parametrize("contrastive", [False, True])
def test_optimizer_overlapped_with_backward_equals_plain_step(use_graphs, contrastive):
    """parallel.StepOverlap: each layer's Adam launch moves behind the backward (side stream); same arithmetic, same
    step counts, same skipped projectors — also when the whole step is a replayed graph (forked capture streams)."""
    steps, kw = (20, {}) if not contrastive else (30, dict(p_drop=0.25, slots=1))
    plain = _run(False, contrastive, steps, fuse=False, **kw)  # (the overlapped optimizer excludes the in-backward one)
    over = _run(use_graphs, contrastive, steps, overlap=True, **kw)
    assert over["trainer"].grad_sync is not None and over["trainer"].grad_sync.opt_stream is not None
    _assert_same_bookkeeping(plain, over)
    assert float((over["losses"] - plain["losses"]).abs().max()) <= 0.05 * float(plain["losses"].abs().max())
    kw = dict(p_drop=0.0, slots=2, lr=5e-4)
    plain, plain2 = _run(False, contrastive, 8, fuse=False, **kw), _run(False, contrastive, 8, fuse=False, **kw)
    _assert_same_numbers(plain, plain2, _run(use_graphs, contrastive, 8, overlap=True, **kw))


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


def test_replayed_adam_reads_the_schedulers_current_hyper_parameters():
    """OneCycleLR changes lr and beta1 every batch; a replayed graph must use the values of THAT step: the device
    hyper-parameter block of every optimizer run is checked against torch's own bias-correction arithmetic."""
    import math

    torch.manual_seed(2)
    model = FmriEncoder(SMALL_DIMS, 200, 25, FmriEncoderConfig(n_subjects=3), **SMALL)
    module = BrainModule(model=model, loss=torch.nn.MSELoss(), optim_config=None, metrics={}, max_epochs=1)
    opt, sched = default_optimizer(model.parameters(), total_steps=20, lr=2e-3, model=model)
    trainer = MiniTrainer(module, opt, sched, use_graphs=True)
    spec = tuple((k, v[0], v[1]) for k, v in SMALL_DIMS.items())
    host = synthetic_batch(batch_size=3, t=74, t_out=25, n_outputs=200, n_subjects=3, seed=1, dims=spec)
    dev = SegmentData(data={k: v.cuda() for k, v in host.data.items()}, segments=host.segments)
    seen = []
    orig = opt.prepare_replay

    def spy(runs):
        group = opt.param_groups[0]
        k = int(opt.state[runs[0]["params"][0]]["step"].item()) + 1
        seen.append((float(group["lr"]), tuple(float(b) for b in group["betas"]), float(group["eps"]), k, [r["lo"] for r in runs]))
        return orig(runs)

    opt.prepare_replay = spy
    for _ in range(7):
        trainer.train_step(dev)
    torch.cuda.synchronize()
    assert len(seen) >= 5 and len({s[0] for s in seen}) == len(seen) and len({s[1][0] for s in seen}) > 1  # lr and beta1 moved every step
    lr, (b1, b2), eps, k, los = seen[-1]
    flat = model._engine.flat
    for lo in los:
        got = flat.adam_hyper[flat.adam_slot[lo]].cpu().tolist()
        want = [b1, b2, lr / (1.0 - b1 ** k), 1.0 / math.sqrt(1.0 - b2 ** k), eps, 0.0]
        for g_, w_ in zip(got[:6], want):
            assert abs(g_ - w_) <= 1e-6 * max(1.0, abs(w_)), (got, want)


GEMM_WEIGHTS = ("to_q.weight", "to_k.weight", "to_v.weight", "to_out.weight", "ff.0.0.weight", "ff.2.weight", "predictor.weights")


def _is_gemm_weight(name):
    return name.endswith(GEMM_WEIGHTS) or (name.startswith("projectors.") and name.endswith(".weight"))


@pytest.mark.parametrize("use_graphs", [False, True])
@pytest.mark.parametrize("contrastive", [False, True])
def test_adam_fused_into_the_wgrad_epilogues_equals_the_plain_step(use_graphs, contrastive):
    """optim.TribeAdam.fuse_backward: the Adam update of every GEMM weight runs inside its weight-gradient GEMM epilogue
    (last backward pass of the step).  Those parameters keep ``grad is None``; everything else — step counts, skipped
    projectors under modality dropout, RNG draws, lr schedule — is the plain step's.  After ONE step (no atomics-order
    noise can have reached a weight gradient yet) parameters, moments and bf16 shadow weights are identical (see TINY)."""
    one_plain = _run(False, contrastive, 1, p_drop=0.0, fuse=False)
    one_fused = _run(False, contrastive, 1, p_drop=0.0, fuse=True)
    fused_names = [n for n in one_fused["none_grads"][0] if n not in one_plain["none_grads"][0]]
    assert fused_names and all(_is_gemm_weight(n) for n in fused_names), fused_names
    assert {n for n in one_plain["state"] if _is_gemm_weight(n)} == set(fused_names)  # every GEMM weight took the fused path
    # the yardstick is a second plain run: bit-identical unless the step itself is not (the InfoNCE gradient of the
    # contrastive branch accumulates with fp32 atomics, so its weight gradients carry run-to-run rounding noise)
    one_plain2 = _run(False, contrastive, 1, p_drop=0.0, fuse=False)
    deterministic = not contrastive  # (two contrastive runs may agree by luck; exact equality is only expected without it)

    def rel(a, b):
        return float((a.float() - b.float()).norm()) / (float(b.float().norm()) + 1e-20)

    # Without the contrastive branch one step is bit-reproducible in practice (relative deviation exactly 0 in every run but
    # one of ~15 full-suite runs on B200, where a single comparison differed; the bound below is 1e-6 of the tensor's
    # norm — three orders of magnitude below one bf16 rounding — so that a lone reordered fp32 atomic cannot fail the suite).
    TINY = 1e-6
    assert contrastive or all(rel(one_plain2["state"][n], one_plain["state"][n]) <= TINY for n in fused_names)
    for n in fused_names:
        assert one_plain["steps_of"][n] == one_fused["steps_of"][n] == 1
        pairs = [(one_plain["state"][n], one_plain2["state"][n], one_fused["state"][n])]
        pairs += [(one_plain["moments"][n][i], one_plain2["moments"][n][i], one_fused["moments"][n][i]) for i in (0, 1)]
        for ref, ref2, got in pairs:
            if deterministic:
                assert rel(got, ref) <= TINY, (n, rel(got, ref))
            else:
                noise = rel(ref2, ref)
                assert rel(got, ref) <= max(10.0 * noise, 1e-5), n
    if deterministic:
        assert rel(one_fused["shadow"], one_plain["shadow"]) <= TINY
        assert abs(float(one_plain["losses"][0]) - float(one_fused["losses"][0])) <= TINY * abs(float(one_plain["losses"][0]))

    # under modality dropout: same discrete bookkeeping apart from the fused weights' missing .grad
    steps, kw = (20, {}) if not contrastive else (30, dict(p_drop=0.25, slots=1))
    plain = _run(False, contrastive, steps, fuse=False, **kw)
    fused = _run(use_graphs, contrastive, steps, fuse=True, **kw)
    if use_graphs:
        assert fused["trainer"]._graphed.replays >= steps // 4
    for a, b in zip(plain["none_grads"], fused["none_grads"]):
        assert set(a) <= set(b) and all(_is_gemm_weight(n) for n in set(b) - set(a))
    fused["none_grads"] = plain["none_grads"]
    _assert_same_bookkeeping(plain, fused)
    assert float((fused["losses"] - plain["losses"]).abs().max()) <= 0.05 * float(plain["losses"].abs().max())
    kw = dict(p_drop=0.0, slots=2, lr=5e-4)
    plain, plain2 = _run(False, contrastive, 8, fuse=False, **kw), _run(False, contrastive, 8, fuse=False, **kw)
    _assert_same_numbers(plain, plain2, _run(use_graphs, contrastive, 8, fuse=True, **kw))
