"""Oracle restatement (oracle/) vs the golden vectors produced by the REFERENCE's own code (oracle/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import tribe_oracle as O

FEATURE_DIMS = {"text": (2, 3072), "audio": (2, 1024), "video": (2, 1408)}


@pytest.fixture(scope="module")
def small(golden_dir):
    return np.load(os.path.join(golden_dir, "small_ops.npz"))


def test_pool_windows_closed_form():
    wins = O.adaptive_pool_windows(298, 100)
    assert wins[:4] == [(0, 3), (2, 6), (5, 9), (8, 12)]
    lens = [e - s for s, e in wins]
    assert lens.count(4) == 96 and lens.count(3) == 4 and wins[-1][1] == 298


@pytest.mark.parametrize("t_in,t_out", [(298, 100), (300, 100), (97, 100), (250, 7)])
def test_pool_matches_reference(small, t_in, t_out):
    x = torch.from_numpy(small[f"pool_{t_in}_{t_out}_x"])
    np.testing.assert_allclose(O.adaptive_avg_pool1d(x, t_out).numpy(), small[f"pool_{t_in}_{t_out}_y"], rtol=1e-6, atol=1e-6)


def test_subject_layers_matches_reference(small):
    out = O.subject_layers(torch.from_numpy(small["sl_x"]), torch.from_numpy(small["sl_subjects"]),
                           torch.from_numpy(small["sl_weights"]), torch.from_numpy(small["sl_bias"]))
    np.testing.assert_allclose(out.numpy(), small["sl_out"], rtol=1e-5, atol=1e-6)


def test_subject_layers_rejects_bad_subject(small):
    with pytest.raises(AssertionError):
        O.subject_layers(torch.from_numpy(small["sl_x"]), torch.full((6, 1), 4),
                         torch.from_numpy(small["sl_weights"]), torch.from_numpy(small["sl_bias"]))


def test_losses_match_reference(small):
    p, t = torch.from_numpy(small["loss_pred"]), torch.from_numpy(small["loss_true"])
    np.testing.assert_allclose(O.pearson_loss(p, t).numpy(), small["pearson_loss"], rtol=1e-6)
    np.testing.assert_allclose(torch.nn.functional.mse_loss(p, t).numpy(), small["mse_loss"], rtol=1e-6)
    q, k = torch.from_numpy(small["nce_q"]), torch.from_numpy(small["nce_k"])
    np.testing.assert_allclose(O.OracleFmriEncoder.info_nce(q, k, 0.07).numpy(), small["nce_loss"], rtol=1e-6)


def test_pearson_variants_agree(small):
    p, t = small["loss_pred"], small["loss_true"]
    # scipy loop (main.py:474-476) on (b d t)-shaped inputs: fold (64,17) as b=64, d=17, t=1
    r = O.multidim_pearson_scipy(p[:, :, None], t[:, :, None])
    np.testing.assert_allclose(r, small["scipy_r"], atol=1e-6)
    np.testing.assert_allclose(O.pearson_columns_f64(p, t), small["scipy_r"], atol=1e-6)
    chunks = [slice(0, 10), slice(10, 45), slice(45, 64)]
    r_stream, r_mean = O.streaming_pearson([torch.from_numpy(p[c]) for c in chunks],
                                           [torch.from_numpy(t[c]) for c in chunks], 17)
    np.testing.assert_allclose(r_stream.numpy(), small["scipy_r"], atol=1e-5)
    np.testing.assert_allclose(r_mean.numpy(), small["metric_pearson_mean"], atol=1e-6)


def test_ensemble_average_weights_sum_to_one():
    rng = np.random.default_rng(0)
    preds, r = rng.normal(size=(3, 11, 5)), rng.uniform(-0.1, 0.4, size=(3, 5))
    out = O.ensemble_average(preds, r, 0.3)
    w = np.exp(r / 0.3) / np.exp(r / 0.3).sum(0)
    np.testing.assert_allclose(out, np.einsum("no,nto->to", w, preds), rtol=1e-12)


@pytest.mark.slow
def test_full_model_matches_reference(golden_dir):
    """Same seed -> same weights (init order), same eval output, same dropout masks / RNG consumption, same
    train-step loss and gradients as the reference's own FmriEncoder + BrainModule."""
    g = np.load(os.path.join(golden_dir, "full_model.npz"))
    meta = json.load(open(os.path.join(golden_dir, "full_model.json")))
    torch.manual_seed(33)
    np.random.seed(33)
    cfg = O.OracleConfig(n_subjects=4, modality_dropout=0.3)
    model = O.OracleFmriEncoder(FEATURE_DIMS, 1000, 100, cfg)
    sd = model.reference_state_dict()
    for name, val in zip(g["param_checksum_names"], g["param_checksum_vals"]):
        assert abs(sd[str(name)].double().sum().item() - val) <= 1e-9 * max(1.0, abs(val)), name

    tiny = O.synthetic_batch(batch_size=1, t=4, t_out=2, seed=99)
    model.train()
    for key, want in meta["dropout_masks"].items():
        if key.endswith("next_rand"):
            continue
        p, seed = key.split(":")
        model.config.modality_dropout = float(p)
        torch.manual_seed(int(seed)), np.random.seed(int(seed))
        with torch.no_grad():
            model.aggregate_features(tiny)
        assert sorted(model.last_dropped) == sorted(want), key
        assert torch.rand(1).item() == meta["dropout_masks"][key + ":next_rand"]
    model.config.modality_dropout = 0.3

    batch = O.synthetic_batch(batch_size=2, seed=1234)
    model.eval()
    torch.manual_seed(123)
    with torch.no_grad():
        y = model(batch)
        assert torch.rand(1).item() == meta["eval_next_rand"]
        y_np = model(batch, pool_outputs=False)
    np.testing.assert_allclose(y.numpy(), g["eval_y"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(y_np[:, ::50, :].numpy(), g["eval_y_nopool_sub"], rtol=1e-4, atol=2e-5)
    loss = torch.nn.functional.mse_loss(O.flatten_bdt(y), O.flatten_bdt(batch.data["fmri"]))
    assert abs(loss.item() - meta["val_loss"]) < 1e-5 * meta["val_loss"]
    r, r_mean = O.streaming_pearson([O.flatten_bdt(y)], [O.flatten_bdt(batch.data["fmri"])], 1000)
    assert abs(r_mean.item() - meta["val_pearson"]) < 1e-6

    model.train()
    torch.manual_seed(meta["train_seed"]), np.random.seed(meta["train_seed"])
    loss, *_ = O.run_step(model, batch)
    loss.backward()
    assert model.last_dropped == meta["train_dropped"]
    assert abs(loss.item() - meta["train_loss"]) < 1e-5 * meta["train_loss"]
    grads = {k.replace("predictor_weights", "predictor.weights").replace("predictor_bias", "predictor.bias"): v.grad
             for k, v in model.named_parameters()}
    assert sorted(k for k, v in grads.items() if v is None) == sorted(meta["grad_none"])
    for name, norm in zip(g["grad_names"], g["grad_norms"]):
        got = grads[str(name)].double().norm().item()
        assert abs(got - norm) <= 2e-4 * max(norm, 1e-12), (name, got, norm)
    np.testing.assert_allclose(grads["predictor.bias"].numpy(), g["grad_predictor_bias"], rtol=1e-3, atol=1e-8)


def test_xt_v127_restatement_is_the_v2_encoder_under_a_head_dim_permutation():
    """Independent check of the 1.27.x restatement (oracle/xt_encoder.py): half-split rotary pairs (i, i + rot/2) are the
    interleaved pairs (2i, 2i+1) after permuting the rotary dims of every q / k head, and attention scores are invariant
    to a common permutation of q and k head dims — so permuting the rows of to_q / to_k turns a v1.27 attention block
    into the v2 one, output for output.  ScaleNorm: g_v127 = g_v2 * sqrt(dim) (away from the eps clamps)."""
    import torch

    from oracle import xt_encoder as X

    torch.manual_seed(0)
    dim, heads, dh, T = 256, 4, 64, 19
    rot = max(dh // 2, 32)
    a1, a2 = X.Attention(dim, dh, heads, "v1.27"), X.Attention(dim, dh, heads, "v2")
    a2.load_state_dict(a1.state_dict())
    perm = torch.arange(dh)
    perm[:rot] = torch.stack((torch.arange(rot // 2), torch.arange(rot // 2) + rot // 2), dim=-1).reshape(-1)  # interleaved slot -> half-split index
    rows = (torch.arange(heads)[:, None] * dh + perm[None, :]).reshape(-1)
    with torch.no_grad():
        a2.to_q.weight.copy_(a1.to_q.weight[rows]), a2.to_k.weight.copy_(a1.to_k.weight[rows])
    x = torch.randn(2, T, dim)
    f1, f2 = X.RotaryEmbedding(rot, semantics="v1.27")(T), X.RotaryEmbedding(rot, semantics="v2")(T)
    torch.testing.assert_close(a1(x, f1), a2(x, f2), rtol=1e-4, atol=1e-5)
    n1, n2 = X.ScaleNorm(dim, "v1.27"), X.ScaleNorm(dim, "v2")
    with torch.no_grad():
        n1.g.fill_(0.37), n2.g.fill_(0.37 / dim ** 0.5)
    torch.testing.assert_close(n1(x), n2(x), rtol=1e-5, atol=1e-6)
    assert abs(float(X.ScaleNorm(dim, "v1.27").g) - dim ** -0.5) < 1e-9 and float(X.ScaleNorm(dim, "v2").g) == 1.0
