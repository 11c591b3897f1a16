"""End-to-end parity of the B200 FmriEncoder / BrainModule against (a) the golden vectors produced by the REFERENCE's
own model.py / pl_module.py (tests/golden, oracle/make_golden.py) and (b) the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): dropout masks / subject gathers / RNG consumption bit-exact; predictions within
1e-2 (bf16 tensor-core operands, fp32 accumulate) measured as max|err| <= 1e-2 * max|ref| and relative L2 <= 1e-2;
Pearson r within 1e-3 absolute."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import algonauts2025_b200  # noqa: E402,F401
from algonauts2025_b200.losses import mse_loss  # noqa: E402
from algonauts2025_b200.metrics import GroupedMetric, MultidimPearsonCorrCoef, compute_multidim_pearson  # noqa: E402
from algonauts2025_b200.model import FmriEncoder, FmriEncoderConfig  # noqa: E402
from algonauts2025_b200.pl_module import BrainModule  # noqa: E402
from algonauts2025_b200.segment import SegmentData, synthetic_batch  # noqa: E402
from oracle import tribe_oracle as O  # noqa: E402

FEATURE_DIMS = {"text": (2, 3072), "audio": (2, 1024), "video": (2, 1408)}
SMALL_DIMS = {"text": (2, 96), "audio": (2, 40), "video": (1, 72)}
SMALL = dict(hidden=384, depth=2, heads=6)  # head_dim 64, rotary dim 32, 128 columns per modality


def rel_l2(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def assert_pred_close(y, ref, tol=1e-2):
    y, ref = y.float().cpu(), ref.float().cpu()
    assert torch.isfinite(y).all()
    assert float((y - ref).abs().max()) <= tol * float(ref.abs().max()), (float((y - ref).abs().max()), float(ref.abs().max()))
    assert rel_l2(y, ref) <= tol, rel_l2(y, ref)


def small_pair(cfg_kw=None, dims=SMALL_DIMS, n_out=200, t_out=25, seed=5):
    """Product model + oracle model with identical weights (state_dict names are shared)."""
    cfg_kw = dict(n_subjects=3, **(cfg_kw or {}))
    torch.manual_seed(seed)
    model = FmriEncoder(dims, n_out, t_out, FmriEncoderConfig(**cfg_kw), **SMALL)
    oracle = O.OracleFmriEncoder(dims, n_out, t_out, O.OracleConfig(**cfg_kw), **SMALL)
    oracle.load_reference_state_dict({k: v.detach().cpu().clone() for k, v in model.state_dict().items()})
    return model, oracle


def small_batch(dims=SMALL_DIMS, b=3, t=74, t_out=25, n_out=200, seed=7, dtype=torch.float32):
    spec = tuple((k, v[0], v[1]) for k, v in dims.items() if v is not None)
    return synthetic_batch(batch_size=b, t=t, t_out=t_out, n_outputs=n_out, n_subjects=3, seed=seed, dims=spec, dtype=dtype)


def as_oracle_batch(batch):
    return O.SegmentData(data=batch.data, segments=batch.segments)


# ----------------------------------------------------------------------------------------------- small model vs oracle
@pytest.mark.parametrize("cfg_kw", [{}, {"layer_aggregation": "mean"}, {"feature_aggregation": "sum"}])
def test_small_forward_matches_oracle(cfg_kw):
    dims = SMALL_DIMS if cfg_kw.get("layer_aggregation") != "mean" else {"text": (2, 96), "audio": (2, 40), "video": (3, 72)}
    model, oracle = small_pair(cfg_kw, dims=dims)
    batch = small_batch(dims)
    model.eval(), oracle.eval()
    with torch.no_grad():
        y = model(batch)
        ref = oracle(as_oracle_batch(batch))
        y_np = model(batch, pool_outputs=False)
        ref_np = oracle(as_oracle_batch(batch), pool_outputs=False)
    assert y.shape == (3, 200, 25) and y.dtype == torch.float32 and y.is_cuda
    assert_pred_close(y, ref)
    assert_pred_close(y_np, ref_np)


def test_small_missing_modality_and_float64_features():
    dims = {"text": (2, 96), "audio": None, "video": (1, 72)}
    model, oracle = small_pair(dims=dims)
    batch = small_batch({"text": (2, 96), "video": (1, 72)}, dtype=torch.float64)
    model.eval(), oracle.eval()
    with torch.no_grad():
        assert_pred_close(model(batch), oracle(as_oracle_batch(batch)))
        agg, agg_ref = model.aggregate_features(batch), oracle.aggregate_features(as_oracle_batch(batch))
    assert_pred_close(agg, agg_ref)
    assert float(agg[:, :, 128:256].abs().max()) == 0.0  # projector-less modality block is exactly zero (model.py:143-144)


def test_small_train_step_gradients_match_oracle():
    """Modality dropout active (same CPU-RNG draws on both sides), MSE loss, every parameter gradient compared."""
    model, oracle = small_pair({"modality_dropout": 0.5})
    batch = small_batch()
    model.train(), oracle.train()
    seed = next(s for s in range(64) if _n_dropped(s, 0.5) == 1)
    torch.manual_seed(seed), np.random.seed(seed)
    loss = mse_loss(model(batch), batch.data["fmri"])
    loss.backward()
    torch.manual_seed(seed), np.random.seed(seed)
    ref_loss, *_ = O.run_step(oracle, as_oracle_batch(batch))
    ref_loss.backward()
    assert model.last_dropped == oracle.last_dropped and len(model.last_dropped) == 1
    assert abs(loss.item() - ref_loss.item()) <= 1e-2 * abs(ref_loss.item())
    ref_grads = {k.replace("predictor_weights", "predictor.weights").replace("predictor_bias", "predictor.bias"): v.grad
                 for k, v in oracle.named_parameters()}
    for name, p in model.named_parameters():
        rg = ref_grads[name]
        if rg is None:
            assert p.grad is None, f"{name}: reference leaves grad=None for a dropped modality's projector"
            continue
        assert p.grad is not None, name
        err = rel_l2(p.grad, rg)
        assert err <= 3e-2, (name, err)


def _n_dropped(seed, p):
    torch.manual_seed(seed), np.random.seed(seed)
    return len(O.draw_modality_dropout(["text", "audio", "video"], p, True))


def test_dropout_masks_and_rng_consumption_bit_exact(golden_dir):
    """The reference's own masks per (p, seed) and the generator position afterwards (tests/golden/full_model.json)."""
    meta = json.load(open(os.path.join(golden_dir, "full_model.json")))
    model, _ = small_pair()
    model.train()
    for key, want in meta["dropout_masks"].items():
        if key.endswith("next_rand"):
            continue
        p, seed = key.split(":")
        model.config.modality_dropout = float(p)
        torch.manual_seed(int(seed)), np.random.seed(int(seed))
        got = model._draw_dropout()
        assert sorted(got) == sorted(want), key
        assert torch.rand(1).item() == meta["dropout_masks"][key + ":next_rand"], key
    model.eval()
    model.config.modality_dropout = 0.3
    batch = small_batch()
    torch.manual_seed(123)
    with torch.no_grad():
        model(batch)
    assert torch.rand(1).item() == meta["eval_next_rand"]  # three draws consumed in eval too (model.py:135-137)


def test_subject_range_assert_and_no_cpu_fallback():
    model, _ = small_pair()
    batch = small_batch()
    bad = SegmentData(data={**batch.data, "subject_id": torch.full((3, 1), 3)}, segments=batch.segments)
    with pytest.raises(AssertionError):
        model(bad)
    with pytest.raises(algonauts2025_b200.TribeError):
        model.pooler(torch.zeros(2, 3, 10))
    # deferred mode (training loops: no host sync inside the step): same check, raised at the next call / flush
    model.defer_subject_check = True
    with torch.no_grad():
        model(bad)
        with pytest.raises(AssertionError):
            model.flush_subject_check()
        model(bad)
        with pytest.raises(AssertionError):
            model(batch)
        model(batch)
        model.flush_subject_check()


def test_transformer_forward_and_state_dict_roundtrip():
    model, oracle = small_pair()
    x = torch.randn(2, 50, 384)
    model.eval(), oracle.eval()
    with torch.no_grad():
        assert_pred_close(model.transformer_forward(x.cuda()), oracle.transformer_forward(x))
    sd = model.state_dict()
    torch.manual_seed(99)
    other = FmriEncoder(SMALL_DIMS, 200, 25, FmriEncoderConfig(n_subjects=3), **SMALL)
    other.load_state_dict(sd)
    batch = small_batch()
    other.eval()
    with torch.no_grad():
        assert torch.equal(other(batch), model(batch))


def test_two_steps_of_adam_track_the_oracle():
    """zero_grad(set_to_none) -> step -> backward -> Adam.step twice: bf16 shadow weights are refreshed after the
    in-place update and the second loss follows the oracle's trajectory."""
    model, oracle = small_pair()
    batch = small_batch()
    model.train(), oracle.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)  # raw-pointer update: shadow refresh must still fire
    opt_ref = torch.optim.Adam(oracle.parameters(), lr=1e-3)
    losses, ref_losses = [], []
    for _ in range(3):
        opt.zero_grad(set_to_none=True), opt_ref.zero_grad(set_to_none=True)
        torch.manual_seed(0)
        loss = mse_loss(model(batch), batch.data["fmri"])
        loss.backward(), opt.step()
        torch.manual_seed(0)
        ref, *_ = O.run_step(oracle, as_oracle_batch(batch))
        ref.backward(), opt_ref.step()
        losses.append(loss.item()), ref_losses.append(ref.item())
    assert losses[2] < losses[0]
    np.testing.assert_allclose(losses, ref_losses, rtol=2e-2)
    flat = model._engine.flat
    flat.refresh_bf16()
    for name, p in model.named_parameters():
        assert torch.equal(flat.view16(name), p.detach().to(torch.bfloat16)), f"stale bf16 shadow for {name}"


# ----------------------------------------------------------------------------------------------- metrics / BrainModule
def test_brain_module_steps_and_metrics():
    model, oracle = small_pair()
    batch = small_batch()
    metrics = torch.nn.ModuleDict({"val/pearson": MultidimPearsonCorrCoef(num_outputs=200),
                                   "val/subj_pearson": GroupedMetric("MultidimPearsonCorrCoef", {"num_outputs": 200})})
    module = BrainModule(model=model, loss=torch.nn.MSELoss(), optim_config=None, metrics=metrics, max_epochs=1)
    module.eval(), oracle.eval()
    with torch.no_grad():
        y_pred, y_true = module.validation_step(batch, 0)
        ref = oracle(as_oracle_batch(batch))
    module.on_validation_epoch_end()
    assert y_pred.device.type == "cpu" and y_true.device.type == "cpu"  # pl_module.py:107
    assert_pred_close(y_pred, ref)
    ref_loss = torch.nn.functional.mse_loss(ref, batch.data["fmri"])
    assert abs(float(module.logged["val/loss"]) - ref_loss.item()) <= 1e-2 * ref_loss.item()
    # metric values vs the oracle's torchmetrics restatement fed with the PRODUCT's predictions (isolates the metric)
    r, r_mean = O.streaming_pearson([O.flatten_bdt(y_pred)], [O.flatten_bdt(y_true)], 200)
    assert abs(float(metrics["val/pearson"].compute()) - r_mean.item()) < 1e-5
    per_parcel = metrics["val/pearson"].per_output().cpu().numpy()
    np.testing.assert_allclose(per_parcel, O.multidim_pearson_scipy(y_pred.numpy(), y_true.numpy()), atol=1e-5)
    subj = batch.data["subject_id"].flatten()
    for key, val in module.logged.items():
        if key.startswith("val/subj_pearson/"):
            s = int(key.rsplit("/", 1)[1])
            sel = subj == s
            want = O.pearson_columns_f64(O.flatten_bdt(y_pred[sel]).numpy(), O.flatten_bdt(y_true[sel]).numpy()).mean()
            assert abs(val - want) < 1e-5
    assert {int(k.rsplit("/", 1)[1]) for k in module.logged if k.startswith("val/subj_pearson/")} == set(subj.tolist())
    # training_step returns a scalar loss with a graph; grads appear on the model
    module.train()
    loss = module.training_step(batch, 0)
    assert loss.dim() == 0 and loss.requires_grad
    loss.backward()
    assert model.predictor.weights.grad is not None
    # final evaluation entry (main.py:459-477)
    r_eval = compute_multidim_pearson(module, [batch, batch])
    module.eval()
    with torch.no_grad():
        yp = module(batch).cpu()
    want = O.multidim_pearson_scipy(torch.cat([yp, yp]).numpy(), torch.cat([batch.data["fmri"]] * 2).numpy())
    np.testing.assert_allclose(r_eval, want, atol=1e-3)


# ----------------------------------------------------------------------------------------------- full-size vs reference
@pytest.fixture(scope="module")
def full_model():
    torch.manual_seed(33)
    np.random.seed(33)
    cfg = FmriEncoderConfig(n_subjects=4, modality_dropout=0.3, feature_aggregation="cat", layer_aggregation="cat")
    return cfg.build(feature_dims=FEATURE_DIMS, n_outputs=1000, n_output_timesteps=100)


def test_full_model_init_matches_reference(full_model, golden_dir):
    g = np.load(os.path.join(golden_dir, "full_model.npz"))
    sd = dict(full_model.named_parameters())
    assert [str(n) for n in g["param_checksum_names"]] == [n for n, _ in full_model.named_parameters()]
    for name, val in zip(g["param_checksum_names"], g["param_checksum_vals"]):
        got = sd[str(name)].detach().double().sum().item()
        assert abs(got - val) <= 1e-6 * max(1.0, abs(val)), name
    assert sum(p.numel() for p in full_model.parameters()) == 932_854_705


def test_full_model_eval_matches_reference_output(full_model, golden_dir):
    """B=2 windows of the full TRIBE shape, eval mode: the reference's own (2, 1000, 100) fp32 output."""
    g = np.load(os.path.join(golden_dir, "full_model.npz"))
    meta = json.load(open(os.path.join(golden_dir, "full_model.json")))
    batch = synthetic_batch(batch_size=2, seed=1234)
    full_model.eval()
    torch.manual_seed(123)
    with torch.no_grad():
        y = full_model(batch)
        assert torch.rand(1).item() == meta["eval_next_rand"]
        y_np = full_model(batch, pool_outputs=False)
        agg = full_model.aggregate_features(batch)
    assert_pred_close(y, torch.from_numpy(g["eval_y"]))
    assert_pred_close(y_np[:, ::50, :], torch.from_numpy(g["eval_y_nopool_sub"]))
    assert_pred_close(agg[:, ::37, ::101], torch.from_numpy(g["eval_agg_sub"]))
    loss = mse_loss(y, batch.data["fmri"])
    assert abs(loss.item() - meta["val_loss"]) <= 1e-2 * meta["val_loss"]
    metric = MultidimPearsonCorrCoef(num_outputs=1000)
    metric.update_bdt(y, batch.data["fmri"].cuda())
    assert abs(float(metric.compute()) - meta["val_pearson"]) <= 1e-3


def test_full_model_train_step_matches_reference_gradients(full_model, golden_dir):
    g = np.load(os.path.join(golden_dir, "full_model.npz"))
    meta = json.load(open(os.path.join(golden_dir, "full_model.json")))
    batch = synthetic_batch(batch_size=2, seed=1234)
    module = BrainModule(model=full_model, loss=torch.nn.MSELoss(), optim_config=None, metrics={}, max_epochs=1)
    module.train()
    for p in full_model.parameters():
        p.grad = None
    torch.manual_seed(meta["train_seed"]), np.random.seed(meta["train_seed"])
    loss = module.training_step(batch, 0)
    loss.backward()
    assert full_model.last_dropped == meta["train_dropped"]
    assert abs(loss.item() - meta["train_loss"]) <= 1e-2 * meta["train_loss"]
    grads = {k: v.grad for k, v in full_model.named_parameters()}
    assert sorted(k for k, v in grads.items() if v is None) == sorted(meta["grad_none"])
    worst = 0.0
    for name, norm in zip(g["grad_names"], g["grad_norms"]):
        got = grads[str(name)].double().norm().item()
        worst = max(worst, abs(got - norm) / max(norm, 1e-12))
        assert abs(got - norm) <= 5e-2 * max(norm, 1e-12), (str(name), got, norm)
    assert rel_l2(grads["predictor.bias"], torch.from_numpy(g["grad_predictor_bias"])) <= 2e-2
    assert rel_l2(grads["time_pos_embed"][0, :298:13, ::97], torch.from_numpy(g["grad_time_pos_embed_sub"])) <= 5e-2
    assert rel_l2(grads["encoder.layers.0.1.to_q.weight"][::193, ::211], torch.from_numpy(g["grad_l0_toq_sub"])) <= 5e-2
    assert rel_l2(grads["encoder.layers.15.1.ff.2.weight"][::193, ::811], torch.from_numpy(g["grad_l15_ff2_sub"])) <= 5e-2
    assert rel_l2(grads["encoder.layers.7.2.residual_scale"], torch.from_numpy(g["grad_residual_scale_l7"])) <= 5e-2


# ----------------------------------------------------------------------------------------------- BASELINE.json config 1
@pytest.mark.parametrize("subject_embedding", [False, True])
def test_config1_single_subject_short_clips_match_oracle(subject_embedding):
    """``grids/test_run``-shaped case (BASELINE.json configs[0]): ONE subject, short clips (T = 61 feature steps -> 20 TRs),
    1000 parcels, tiny batch — forward, loss and readout gradient against the CPU oracle; the subject gather degenerates
    to a single weight slice and the pooling windows are ragged (61 -> 20)."""
    dims = {"text": (2, 96), "audio": (2, 40), "video": (1, 72)}
    cfg_kw = dict(n_subjects=1, subject_embedding=subject_embedding)
    torch.manual_seed(9)
    model = FmriEncoder(dims, 1000, 20, FmriEncoderConfig(**cfg_kw), **SMALL)
    oracle = O.OracleFmriEncoder(dims, 1000, 20, O.OracleConfig(**cfg_kw), **SMALL)
    oracle.load_reference_state_dict({k: v.detach().cpu().clone() for k, v in model.state_dict().items()})
    spec = tuple((k, v[0], v[1]) for k, v in dims.items())
    batch = synthetic_batch(batch_size=2, t=61, t_out=20, n_outputs=1000, n_subjects=1, seed=3, dims=spec)
    assert int(batch.data["subject_id"].max()) == 0
    model.train(), oracle.train()
    y = model(batch)
    loss = mse_loss(y, batch.data["fmri"])
    loss.backward()
    ref_loss, ref_y, *_ = O.run_step(oracle, as_oracle_batch(batch))
    ref_loss.backward()
    assert y.shape == (2, 1000, 20)
    assert_pred_close(y.detach(), ref_y.detach())
    assert abs(loss.item() - ref_loss.item()) <= 1e-2 * ref_loss.item()
    got, want = model.predictor.weights.grad.float().cpu(), oracle.predictor_weights.grad
    assert rel_l2(got, want) <= 3e-2
    with torch.no_grad():
        assert_pred_close(model(batch, pool_outputs=False), oracle(as_oracle_batch(batch), pool_outputs=False))


# ----------------------------------------------------------------------------------------------- API holes closed in round 2
def test_transformer_forward_with_subject_embedding_and_gradients():
    """model.py:167-174 with ``subject_embedding=True`` (the ensemble grid samples it, run_ensemble.py:31):
    ``x + time_pos_embed[:, :T] + subject_embed(subject_id)`` -> encoder; forward and d/dx, d/d(subject_embed) vs oracle."""
    model, oracle = small_pair(dict(subject_embedding=True))
    torch.manual_seed(3)
    x = torch.randn(3, 50, 384)
    sid = torch.tensor([[2], [0], [2]])
    model.train(), oracle.train()
    xg = x.cuda().requires_grad_(True)
    y = model.transformer_forward(xg, sid.cuda())
    xo = x.clone().requires_grad_(True)
    yo = oracle.transformer_forward(xo, sid)
    assert_pred_close(y.detach(), yo.detach())
    w = torch.randn_like(yo)
    (y * w.cuda()).sum().backward()
    (yo * w).sum().backward()
    assert rel_l2(xg.grad, xo.grad) < 3e-2
    assert rel_l2(model.subject_embed.weight.grad, oracle.subject_embed.weight.grad) < 3e-2
    assert float(model.subject_embed.weight.grad[1].abs().max()) == 0.0  # subject 1 not in the batch
    with pytest.raises(algonauts2025_b200.TribeError):
        model.transformer_forward(xg)  # subject ids are required once the embedding exists


def test_standalone_subject_layers_is_differentiable():
    """common.py:45-67 imported on its own: output, d/dx, d/dweights (zero for absent subjects), d/dbias vs torch autograd
    on the reference formula ``einsum("bct,bcd->bdt", x, W[subjects]) + bias[subjects]``."""
    from algonauts2025_b200.model import SubjectLayers

    torch.manual_seed(8)
    B, C, T, D, S = 5, 128, 24, 200, 4
    layer = SubjectLayers(C, D, S, bias=True).cuda()
    subj = torch.tensor([[3], [0], [3], [1], [0]])
    x = torch.randn(B, C, T)
    xg = x.cuda().requires_grad_(True)
    out = layer(xg, subj.cuda())
    W = layer.weights.detach().cpu().clone().requires_grad_(True)
    bias = layer.bias.detach().cpu().clone().requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    ref = torch.einsum("bct,bcd->bdt", xr, W.index_select(0, subj.flatten())) + bias.index_select(0, subj.flatten()).view(B, D, 1)
    assert_pred_close(out.detach(), ref.detach())
    g = torch.randn_like(ref)
    (out * g.cuda()).sum().backward()
    (ref * g).sum().backward()
    assert rel_l2(xg.grad, xr.grad) < 2e-2
    assert rel_l2(layer.weights.grad, W.grad) < 2e-2
    assert float(layer.weights.grad[2].abs().max()) == 0.0
    # the bias gradient sums the bf16-rounded output gradient over (samples of the subject, t): entries of magnitude ~6
    # carry ~3e-2 of absolute rounding noise, so compare in relative L2 like every other gradient
    assert rel_l2(layer.bias.grad, bias.grad) < 2e-2
    assert float(layer.bias.grad[2].abs().max()) == 0.0
    with pytest.raises(AssertionError):
        layer(xg, torch.tensor([[4], [0], [0], [0], [0]]).cuda())


# ----------------------------------------------------------------------------------------------- x_transformers 1.27.x semantics
def _pair_v127(cfg_kw=None, seed=5):
    cfg_kw = dict(n_subjects=3, **(cfg_kw or {}))
    torch.manual_seed(seed)
    model = FmriEncoder(SMALL_DIMS, 200, 25, FmriEncoderConfig(xt_semantics="v1.27", **cfg_kw), **SMALL)
    oracle = O.OracleFmriEncoder(SMALL_DIMS, 200, 25, O.OracleConfig(**cfg_kw), xt_semantics="v1.27", **SMALL)
    oracle.load_reference_state_dict({k: v.detach().cpu().clone() for k, v in model.state_dict().items()})
    return model, oracle


def test_xt_semantics_v127_forward_and_gradients_match_the_oracle():
    """modeling_utils/pyproject.toml:12 admits x_transformers 1.27.x, whose ScaleNorm (x / norm.clamp(1e-5) * g with
    g = dim ** -0.5) and rotary embedding (half-split pairs i, i + rot/2) differ from >= 2.x while the state-dict keys
    are identical.  ``xt_semantics="v1.27"``: forward, loss and every parameter gradient vs oracle/xt_encoder.py in that
    mode; the ScaleNorm gains are initialised like that release; and the two semantics really differ on the same weights."""
    model, oracle = _pair_v127()
    H = SMALL["hidden"]
    for g in (model.encoder.final_norm.g, model.encoder.layers[0][0][0].g):
        assert abs(float(g.detach()) - H ** -0.5) < 1e-7  # fp32 of dim ** -0.5
    batch = small_batch()
    model.eval(), oracle.eval()
    with torch.no_grad():
        y, ref = model(batch), oracle(as_oracle_batch(batch))
        lat = model.transformer_forward(torch.randn(2, 50, H, generator=torch.Generator().manual_seed(3)).cuda())
        lat_ref = oracle.transformer_forward(torch.randn(2, 50, H, generator=torch.Generator().manual_seed(3)))
    assert_pred_close(y, ref)
    assert_pred_close(lat, lat_ref)
    # same weights under the other release's arithmetic: a clearly different function
    torch.manual_seed(5)
    other = FmriEncoder(SMALL_DIMS, 200, 25, FmriEncoderConfig(n_subjects=3), **SMALL)
    other.load_state_dict(model.state_dict())
    other.eval()
    with torch.no_grad():
        y2 = other(batch)
    assert rel_l2(y2, ref) > 0.1
    # train step: every gradient
    model.train(), oracle.train()
    torch.manual_seed(11), np.random.seed(11)
    loss = mse_loss(model(batch), batch.data["fmri"])
    loss.backward()
    torch.manual_seed(11), np.random.seed(11)
    ref_loss, *_ = O.run_step(oracle, as_oracle_batch(batch))
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-2 * abs(ref_loss.item())
    ref_grads = {k.replace("predictor_weights", "predictor.weights").replace("predictor_bias", "predictor.bias"): v.grad
                 for k, v in oracle.named_parameters()}
    for name, p in model.named_parameters():
        assert p.grad is not None and ref_grads[name] is not None, name
        err = rel_l2(p.grad, ref_grads[name])
        # the scalar ScaleNorm gains sum signed bf16-rounded terms over every token (cancellation): a little more head-room
        assert err <= (5e-2 if name.endswith(".g") else 3e-2), (name, err)


def test_rope_half_kernel_matches_the_oracle_rotation_and_its_transpose():
    from oracle import xt_encoder as X

    from algonauts2025_b200 import ops

    torch.manual_seed(0)
    rows_b, T, heads, dh = 3, 37, 4, 64
    rot = max(dh // 2, 32)
    x = torch.randn(rows_b * T, 3 * heads * dh)
    xb = x.to(torch.bfloat16).cuda()
    emb = X.RotaryEmbedding(rot, semantics="v1.27")
    freqs = emb(T)
    ang = torch.arange(T).float()[:, None] * emb.inv_freq[None, :]
    table = torch.stack((ang.cos(), ang.sin()), dim=-1).contiguous().cuda()
    ref = xb.float().cpu().view(rows_b, T, 3 * heads, dh).clone()
    q = ref[:, :, : 2 * heads].permute(0, 2, 1, 3)  # (b, h, T, dh)
    ref[:, :, : 2 * heads] = X.apply_rotary(q, freqs, "v1.27").permute(0, 2, 1, 3)
    y = xb.clone()
    ops.rope_half(y, 0, 2 * heads, dh, rot, table, T)
    got = y.float().cpu().view(rows_b, T, 3 * heads, dh)
    assert float((got - ref).abs().max()) <= 2e-2 * float(ref.abs().max())
    assert torch.equal(got[:, :, 2 * heads:], xb.float().cpu().view(rows_b, T, 3 * heads, dh)[:, :, 2 * heads:])  # v heads untouched
    ops.rope_half(y, 0, 2 * heads, dh, rot, table, T, sign=-1.0)  # the transpose is the inverse rotation
    assert float((y.float().cpu() - xb.float().cpu()).abs().max()) <= 3e-2 * float(xb.float().abs().max())
