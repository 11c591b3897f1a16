"""GPU parity of the §8f kernels (through the C ABI) against the reference-generated vectors (tests/golden/aux_ops.npz)
and the CPU oracle (oracle/aux_oracle.py).  Window gathers and ranks are bit-exact; float work: losses 1e-5 relative,
gradients 1e-5 of their scale, ensemble averages 1e-6."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import algonauts2025_b200  # noqa: E402,F401
from algonauts2025_b200 import ensemble, losses as L, ops, windows as W  # noqa: E402
from algonauts2025_b200.metrics import Rank, TopkAcc  # noqa: E402
from algonauts2025_b200.swa import SwaAverager  # noqa: E402
from oracle import aux_oracle as A  # noqa: E402

DEV = "cuda"


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "aux_ops.npz"))


# ------------------------------------------------------------------------------------------------------------ losses
@pytest.mark.parametrize("red", ["mean", "sum"])
def test_pearson_loss_matches_reference_value_and_grad(g, red):
    x = torch.from_numpy(g["loss_pred"]).to(DEV).requires_grad_(True)
    t = torch.from_numpy(g["loss_true"]).to(DEV)
    val = L.PearsonLoss(reduction=red)(x, t)
    (3.0 * val).backward()  # non-trivial upstream gradient (device scalar)
    np.testing.assert_allclose(val.item(), g[f"pearson_{red}"], rtol=1e-5, atol=1e-6)
    ref = 3.0 * g[f"pearson_{red}_grad"]
    np.testing.assert_allclose(x.grad.cpu().numpy(), ref, rtol=1e-4, atol=1e-5 * np.abs(ref).max())


def test_pearson_loss_bdt_layout_equals_flattened_module():
    """pl_module.py:54-56 flattens (B, D, T) -> ((b t), d) before the loss; the fused path indexes parcels in place."""
    torch.manual_seed(3)
    for Bsz, D, T in ((4, 50, 100), (3, 37, 25), (2, 8, 7)):
        true = torch.randn(Bsz, D, T)
        pred = (0.3 * true + torch.randn(Bsz, D, T)).requires_grad_(True)
        ref_val, ref_grad = A.pearson_loss_and_grad(pred.detach().permute(0, 2, 1).reshape(-1, D).numpy(), true.permute(0, 2, 1).reshape(-1, D).numpy())
        pd = pred.detach().to(DEV).requires_grad_(True)
        val = L.fused_loss(L.PearsonLoss(), pd, true.to(DEV))
        val.backward()
        np.testing.assert_allclose(val.item(), ref_val, rtol=1e-5, atol=1e-6)
        got = pd.grad.cpu().permute(0, 2, 1).reshape(-1, D).numpy()
        np.testing.assert_allclose(got, ref_grad, rtol=1e-4, atol=1e-5 * np.abs(ref_grad).max())


@pytest.mark.parametrize("name,mod", [("smooth_l1", torch.nn.SmoothL1Loss()), ("smooth_l1_b05", torch.nn.SmoothL1Loss(beta=0.5)),
                                      ("huber", torch.nn.HuberLoss()), ("huber_d2", torch.nn.HuberLoss(delta=2.0)), ("l1", torch.nn.L1Loss())])
def test_point_losses_match_torch_modules(g, name, mod):
    x = (2.0 * torch.from_numpy(g["loss_pred"])).to(DEV).requires_grad_(True)
    t = torch.from_numpy(g["loss_true"]).to(DEV)
    val = L.fused_loss(mod, x, t)
    assert val is not None
    val.backward()
    np.testing.assert_allclose(val.item(), g[name], rtol=1e-5)
    np.testing.assert_allclose(x.grad.cpu().numpy(), g[f"{name}_grad"], rtol=1e-5, atol=1e-9)
    # odd sizes / unaligned tails
    torch.manual_seed(1)
    a, b = torch.randn(1237, device=DEV) * 2, torch.randn(1237, device=DEV)
    a.requires_grad_(True)
    v = L.fused_loss(mod, a, b)
    v.backward()
    a2 = a.detach().clone().requires_grad_(True)
    ref = mod(a2, b)
    ref.backward()
    torch.testing.assert_close(v, ref, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(a.grad, a2.grad, rtol=1e-5, atol=1e-9)


def test_unfused_losses_fall_back_to_the_module():
    assert L.fused_loss(torch.nn.MSELoss(reduction="sum"), torch.zeros(2, 3, 4, device=DEV), torch.zeros(2, 3, 4, device=DEV)) is None


# ------------------------------------------------------------------------------------------------------------ retrieval
def test_retrieval_ranks_match_reference(g):
    x, y = torch.from_numpy(g["rank_x"]).to(DEV), torch.from_numpy(g["rank_y"]).to(DEV)
    ranks, scores = ops.retrieval_ranks(x, y, want_scores=True)
    np.testing.assert_allclose(scores.cpu().numpy(), g["rank_scores"], rtol=1e-5, atol=1e-6, equal_nan=True)
    m1, m5, mr = TopkAcc(topk=1), TopkAcc(topk=5), Rank(reduction="mean")
    for m in (m1, m5, mr):
        m.update(x, y)
        m.update(x[:7], y[:7])
    np.testing.assert_array_equal(m1.ranks.cpu().numpy(), g["rank_ranks"])  # incl. the tie (0.5 steps) and the NaN query (n // 2)
    assert m1.compute().item() == pytest.approx(float(g["rank_top1"]))
    assert m5.compute().item() == pytest.approx(float(g["rank_top5"]))
    assert mr.compute().item() == pytest.approx(float(g["rank_ranks"].mean()))


def test_retrieval_update_bdt_fuses_the_time_average():
    torch.manual_seed(5)
    true = torch.randn(16, 1000, 100)
    pred = 0.05 * true + torch.randn(16, 1000, 100)
    m = TopkAcc(topk=1)
    m.update_bdt(pred.to(DEV), true.to(DEV))
    ref = A.retrieval_ranks(pred.mean(-1).numpy().astype(np.float64), true.mean(-1).numpy().astype(np.float64))
    np.testing.assert_array_equal(m.ranks.cpu().numpy(), ref)
    torch.testing.assert_close(ops.mean_lastdim(pred.to(DEV)).cpu(), pred.mean(-1), rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------------------------ windows
def test_window_gather_is_bit_exact_against_reference_timed_arrays(g):
    cases = g["win_cases"]
    by_shape = {}
    for i, (n, a0, f, ws, wd) in enumerate(cases):
        by_shape.setdefault((f, wd), []).append(i)
    for (f, wd), idx in by_shape.items():
        store = W.TimelineStore()
        wins = []
        for i in idx:
            n, a0, _, ws, _ = cases[i]
            store.add("feat", f"tl{i}", g[f"win_arr_{i}"], start=a0, frequency=f)
            wins.append((f"tl{i}", ws))
        out = store.assemble("feat", wins, duration=wd).cpu().numpy()   # ONE launch for the whole batch of windows
        for k, i in enumerate(idx):
            np.testing.assert_array_equal(out[k], g[f"win_out_{i}"])


def test_window_batches_feed_the_encoder_shapes():
    """Full-shape batch: 16 windows x (2, 3072, 298) text stacks gathered from 4 timelines + float64 video cache."""
    rng = np.random.default_rng(0)
    store = W.TimelineStore()
    for tl in range(4):
        store.add("text", f"t{tl}", rng.standard_normal((2, 3072, 1300), dtype=np.float32), start=0.0, frequency=2.0)
        store.add("video", f"t{tl}", rng.standard_normal((2, 64, 1300)), start=0.0, frequency=2.0)  # float64 like video.py:230
        store.add("fmri", f"t{tl}", rng.standard_normal((1000, 436), dtype=np.float32), start=0.0, frequency=1 / 1.49)
    wins = [(f"t{i % 4}", float(s)) for i, s in enumerate(np.tile(W.timeline_windows(0.0, 650.0)[0], 4)[:16])]
    batch = store.batch(wins, ["text", "video", "fmri"], subject_ids=[i % 4 for i in range(16)])
    assert batch.data["text"].shape == (16, 2, 3072, 298) and batch.data["video"].shape == (16, 2, 64, 298)
    assert batch.data["fmri"].shape == (16, 1000, 100) and batch.data["subject_id"].shape == (16, 1)
    for k, (tl, ws) in enumerate(wins[:6]):
        for mod, f in (("text", 2.0), ("fmri", 1 / 1.49)):
            arr = store.arrays[(mod, tl)]["data"].cpu().numpy()
            np.testing.assert_array_equal(batch.data[mod][k].cpu().numpy(), A.assemble_window(arr, 0.0, f, ws, 149.0).astype(np.float32))


# ------------------------------------------------------------------------------------------------------------ ensemble / SWA
def test_ensemble_average_matches_reference_average_submissions(g):
    M = g["ens_pearsons"].shape[0]
    members = [{"sub-01": {"c1": g[f"ens_in_{m}_sub-01_c1"], "c2": g[f"ens_in_{m}_sub-01_c2"]}, "sub-02": {"c1": g[f"ens_in_{m}_sub-02_c1"]}}
               for m in range(M)]
    for tag, kw in (("voxel", dict(weigh_by_score=True, per_voxel_weights=True, temperature=0.3)),
                    ("scalar", dict(weigh_by_score=True, per_voxel_weights=False, temperature=0.05)), ("mean", dict(weigh_by_score=False))):
        res = ensemble.average_submissions(members, pearsons=g["ens_pearsons"], scores=g["ens_scores"], **kw)
        for sub, chunks in res.items():
            for c, v in chunks.items():
                np.testing.assert_allclose(v, g[f"ens_{tag}_{sub}_{c}"], rtol=1e-5, atol=1e-6)
    # weights normalised over members instead (a convex combination per parcel)
    w = ensemble.member_weights(g["ens_pearsons"], temperature=0.3, softmax_over="members")
    torch.testing.assert_close(w.sum(0), torch.ones(w.shape[1], device=DEV), rtol=1e-5, atol=1e-6)


def test_ensemble_average_large_and_ragged():
    rng = np.random.default_rng(2)
    for M, N, O in ((8, 2500, 1000), (3, 17, 1003), (50, 9, 64)):
        preds = rng.standard_normal((M, N, O)).astype(np.float32)
        r = rng.uniform(0, 0.4, (M, O)).astype(np.float32)
        w = ensemble.member_weights(r, temperature=0.3)
        got = ensemble.average_predictions(preds, w).cpu().numpy()
        want = A.average_members(preds, r, weigh_by_score=True, per_voxel_weights=True, temperature=0.3)
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(ensemble.average_predictions(preds, None).cpu().numpy(), preds.astype(np.float64).mean(0), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("loss_name", ["PearsonLoss", "SmoothL1Loss", "HuberLoss"])
def test_brain_module_trains_with_the_grid_losses(loss_name):
    """run_ensemble.py:29 samples the loss; BrainModule._run_step (pl_module.py:46-107) must give the oracle's loss value
    (bf16 predictions: 1e-2 relative) and the gradients must flow through the fused loss kernels into the encoder."""
    from algonauts2025_b200.model import FmriEncoder, FmriEncoderConfig
    from algonauts2025_b200.pl_module import BrainModule
    from algonauts2025_b200.segment import synthetic_batch
    from oracle import tribe_oracle as O

    dims = {"text": (2, 96), "audio": (2, 40), "video": (1, 72)}
    small = dict(hidden=384, depth=2, heads=6)
    torch.manual_seed(5)
    model = FmriEncoder(dims, 200, 25, FmriEncoderConfig(n_subjects=3), **small)
    oracle = O.OracleFmriEncoder(dims, 200, 25, O.OracleConfig(n_subjects=3), **small)
    oracle.load_reference_state_dict({k: v.detach().cpu().clone() for k, v in model.state_dict().items()})
    spec = tuple((k, v[0], v[1]) for k, v in dims.items())
    batch = synthetic_batch(batch_size=3, t=74, t_out=25, n_outputs=200, n_subjects=3, seed=7, dims=spec)
    ours = {"PearsonLoss": L.PearsonLoss(), "SmoothL1Loss": torch.nn.SmoothL1Loss(), "HuberLoss": torch.nn.HuberLoss()}[loss_name]
    ref_fn = {"PearsonLoss": O.pearson_loss, "SmoothL1Loss": torch.nn.SmoothL1Loss(), "HuberLoss": torch.nn.HuberLoss()}[loss_name]
    module = BrainModule(model=model, loss=ours, optim_config=None, metrics={"val/retrieval_top1": TopkAcc(topk=1)}, max_epochs=1)
    module.train()
    loss = module.training_step(batch, 0)
    loss.backward()
    oracle.train()
    ref_loss, ref_pred, ref_true, _ = O.run_step(oracle, O.SegmentData(data=batch.data, segments=batch.segments), loss_fn=ref_fn)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-2 * abs(ref_loss.item()) + 1e-4, (loss.item(), ref_loss.item())
    got, want = model.predictor.weights.grad.float().cpu(), oracle.predictor_weights.grad
    assert torch.isfinite(got).all()
    assert float((got - want).norm() / want.norm()) < 3e-2, float((got - want).norm() / want.norm())
    module.eval()
    with torch.no_grad():
        y_pred, y_true = module.validation_step(batch, 0)
    assert y_pred.shape == (3, 200, 25) and not y_pred.is_cuda
    top1 = module.metrics["val/retrieval_top1"].compute().item()
    ref_ranks = A.retrieval_ranks(ref_pred.detach().mean(-1).numpy().astype(np.float64), ref_true.mean(-1).numpy().astype(np.float64))
    assert 0.0 <= top1 <= 1.0 and abs(top1 - A.topk_acc(ref_ranks, 1)) <= 1.0 / 3 + 1e-6  # 3 queries: at most one flip from bf16 noise


def test_swa_running_average_of_the_flat_parameter_buffer():
    from algonauts2025_b200.model import FmriEncoder, FmriEncoderConfig

    torch.manual_seed(0)
    model = FmriEncoder({"text": (2, 96), "audio": (2, 40), "video": (1, 72)}, 50, 25, FmriEncoderConfig(n_subjects=2), hidden=384, depth=2, heads=6)
    swa = SwaAverager(model)
    snaps = []
    for step in range(4):
        with torch.no_grad():
            for p in model.parameters():
                p.add_(0.1 * torch.randn_like(p))
        snaps.append({k: v.detach().clone() for k, v in model.state_dict().items()})
        swa.update_parameters()
    avg = swa.averaged_state_dict()
    for name, _ in model.named_parameters():
        want = torch.stack([s[name] for s in snaps]).mean(0)
        torch.testing.assert_close(avg[name], want, rtol=1e-5, atol=1e-6)
    live = {k: v.detach().clone() for k, v in model.state_dict().items()}
    swa.swap_into_model()
    for name, _ in model.named_parameters():
        torch.testing.assert_close(model.state_dict()[name], torch.stack([s[name] for s in snaps]).mean(0), rtol=1e-5, atol=1e-6)
    swa.swap_into_model()
    for name, _ in model.named_parameters():
        torch.testing.assert_close(model.state_dict()[name], live[name])


# ------------------------------------------------------------------------------------------------------------ submission
def test_transpose_and_flatten_are_bit_exact_with_gradients():
    from algonauts2025_b200.pl_module import _flatten_bdt

    torch.manual_seed(0)
    for shape in ((3, 1000, 100), (2, 37, 298), (1, 5, 1), (4, 33, 65)):
        x = torch.randn(*shape)
        y = ops.transpose_last2(x.to(DEV))
        assert torch.equal(y.cpu(), x.transpose(1, 2).contiguous())
        xd = x.to(DEV).requires_grad_(True)
        flat = _flatten_bdt(xd)
        assert torch.equal(flat.detach().cpu(), x.permute(0, 2, 1).reshape(-1, shape[1]))
        w = torch.randn_like(flat)
        (flat * w).sum().backward()
        assert torch.equal(xd.grad.cpu(), w.cpu().view(shape[0], shape[2], shape[1]).transpose(1, 2).contiguous())


def test_submission_assembler_matches_reference_benchmark_callback(g):
    from algonauts2025_b200.submission import SubmissionAssembler

    layout = [[tuple(x.split("|")) for x in row] for row in g["sub_layout"]]
    samples = {}
    for k in g.files:
        if k.startswith("sub_out_"):
            _, _, subj, chunk = k.split("_")
            samples.setdefault(subj, {})[chunk] = g[k].shape[0]
    asm = SubmissionAssembler()
    for b, row in enumerate(layout):
        asm.add_batch(torch.from_numpy(g[f"sub_pred_{b}"]).to(DEV), [s for s, _ in row], [c for _, c in row])  # raw event labels
    out = asm.finalize(samples)
    assert set(out) == set(samples)
    for subj, per in out.items():
        for chunk, arr in per.items():
            np.testing.assert_array_equal(arr, g[f"sub_out_{subj}_{chunk}"])
    # several windows per chunk (where the reference's float slice index raises): the oracle's overlap-0 assembly
    rng = np.random.default_rng(4)
    batches = [rng.standard_normal((4, 1000, 100)).astype(np.float32) for _ in range(3)]
    labels = [[("sub-01", "s07e01a"), ("sub-02", "s07e01a"), ("sub-01", "s07e01a"), ("sub-01", "s07e01b")],
              [("sub-02", "s07e01a"), ("sub-02", "s07e01b"), ("sub-01", "s07e01b"), ("sub-01", "s07e01a")],
              [("sub-02", "s07e01b"), ("sub-01", "s07e01b"), ("sub-02", "s07e01a"), ("sub-02", "s07e01b")]]
    want_n = {"sub-01": {"s07e01a": 287, "s07e01b": 300}, "sub-02": {"s07e01a": 250, "s07e01b": 299}}
    asm.reset()
    for y, row in zip(batches, labels):
        asm.add_batch(torch.from_numpy(y).to(DEV), [s for s, _ in row], [c for _, c in row])
    got = asm.finalize(want_n)
    ref = A.assemble_submission(batches, labels, want_n)
    for subj, per in ref.items():
        for chunk, arr in per.items():
            np.testing.assert_array_equal(got[subj][chunk], arr)
    with pytest.raises(ValueError):
        asm.finalize({"sub-01": {"s07e01a": 301, "s07e01b": 1}, "sub-02": {"s07e01a": 1, "s07e01b": 1}})
