"""TEST INFRASTRUCTURE — generates ``tests/golden/*`` from the REFERENCE's own code.

Run in the build container only (needs /root/reference):  ``python -m oracle.make_golden``

The reference's ``algonauts2025/model.py``, ``algonauts2025/pl_module.py``,
``modeling_utils/modeling_utils/models/common.py``, ``.../losses/losses.py`` and ``.../metrics/base.py`` are imported
unmodified from /root/reference; the packages that are absent from this image are satisfied by ``oracle/shims``
(``exca``/``lightning``: inert plumbing; ``x_transformers``/``torchmetrics``: the oracle restatements, see their
headers — those two blocks are therefore *not* independently pinned by these vectors).
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle", "shims"), REF, f"{REF}/data_utils", f"{REF}/modeling_utils"]

from algonauts2025.model import FmriEncoder, FmriEncoderConfig  # noqa: E402
from algonauts2025.pl_module import BrainModule  # noqa: E402
from data_utils.dataloader import SegmentData  # noqa: E402
from modeling_utils.losses.losses import PearsonLoss  # noqa: E402
from modeling_utils.metrics.base import GroupedMetric, MultidimPearsonCorrCoef  # noqa: E402
from modeling_utils.models.common import SubjectLayers  # noqa: E402

from oracle.tribe_oracle import synthetic_batch  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
FEATURE_DIMS = {"text": (2, 3072), "audio": (2, 1024), "video": (2, 1408)}


def ref_batch(**kw):
    b = synthetic_batch(**kw)
    return SegmentData(data=b.data, segments=b.segments)


def small_ops():
    out = {}
    g = torch.Generator().manual_seed(5)
    # SubjectLayers (common.py:14-71)
    torch.manual_seed(11)
    sl = SubjectLayers(in_channels=48, out_channels=20, n_subjects=4, bias=True)
    x = torch.randn(6, 48, 13, generator=g)
    subj = torch.tensor([[3], [0], [1], [3], [2], [0]])
    out["sl_weights"], out["sl_bias"] = sl.weights.detach().numpy(), sl.bias.detach().numpy()
    out["sl_x"], out["sl_subjects"] = x.numpy(), subj.numpy()
    out["sl_out"] = sl(x, subj).detach().numpy()
    # AdaptiveAvgPool1d(100) on T=298 and a few other lengths (model.py:60)
    for t_in, t_out in ((298, 100), (300, 100), (97, 100), (250, 7)):
        xp = torch.randn(3, 5, t_in, generator=g)
        out[f"pool_{t_in}_{t_out}_x"] = xp.numpy()
        out[f"pool_{t_in}_{t_out}_y"] = torch.nn.AdaptiveAvgPool1d(t_out)(xp).numpy()
    # PearsonLoss (losses.py:11-42) and nn.MSELoss on (N, O)
    p, t = torch.randn(64, 17, generator=g), torch.randn(64, 17, generator=g)
    t = 0.3 * p + t
    out["loss_pred"], out["loss_true"] = p.numpy(), t.numpy()
    out["pearson_loss"] = PearsonLoss()(p, t).numpy()
    out["mse_loss"] = torch.nn.MSELoss()(p, t).numpy()
    # InfoNCE (model.py:208-221)
    q, k = torch.randn(2, 9, 32, generator=g), torch.randn(2, 9, 32, generator=g)
    out["nce_q"], out["nce_k"] = q.numpy(), k.numpy()
    out["nce_loss"] = FmriEncoder._info_nce(q, k, tau=0.07).numpy()
    # metrics (metrics/base.py:26-29, 39-91) streamed in 3 uneven batches
    mp, gm = MultidimPearsonCorrCoef(num_outputs=17), GroupedMetric("MultidimPearsonCorrCoef", {"num_outputs": 17})
    groups = torch.tensor([0] * 20 + [2] * 30 + [1] * 14)
    for sl_ in (slice(0, 10), slice(10, 45), slice(45, 64)):
        mp.update(p[sl_], t[sl_])
        gm.update(p[sl_], t[sl_], groups=groups[sl_])
    out["metric_groups"] = groups.numpy()
    out["metric_pearson_mean"] = mp.compute().numpy()
    gd = gm.compute()
    out["metric_grouped_keys"] = np.array(list(gd.keys()))
    out["metric_grouped_vals"] = np.array(list(gd.values()), dtype=np.float64)
    from scipy.stats import pearsonr

    out["scipy_r"] = np.array([pearsonr(t[:, j].numpy(), p[:, j].numpy())[0] for j in range(17)], dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, "small_ops.npz"), **out)
    print("small_ops.npz written")


def full_model():
    t0 = time.time()
    torch.manual_seed(33)
    np.random.seed(33)
    cfg = FmriEncoderConfig(n_subjects=4, modality_dropout=0.3, feature_aggregation="cat", layer_aggregation="cat")
    model = cfg.build(feature_dims=FEATURE_DIMS, n_outputs=1000, n_output_timesteps=100)
    print(f"reference model built in {time.time() - t0:.1f}s, params={sum(p.numel() for p in model.parameters())}")
    out, meta = {}, {}
    out["param_checksum_names"] = np.array([k for k, _ in model.named_parameters()])
    out["param_checksum_vals"] = np.array([v.double().sum().item() for _, v in model.named_parameters()])

    # ---- dropout masks (bit-exact target): which modalities the reference zeroes, per (seed, p), train mode ------
    tiny = ref_batch(batch_size=1, t=4, t_out=2, seed=99)
    masks = {}
    model.train()
    for p in (0.3, 0.6, 0.95):
        model.config.modality_dropout = p
        for seed in range(24):
            torch.manual_seed(seed)
            np.random.seed(seed)
            with torch.no_grad():
                agg = model.aggregate_features(tiny)
            blocks = agg.reshape(1, 4, 3, 1024).abs().sum(dim=(0, 1, 3))
            masks[f"{p}:{seed}"] = [m for m, s in zip(FEATURE_DIMS, blocks.tolist()) if s == 0.0]
            masks[f"{p}:{seed}:next_rand"] = torch.rand(1).item()
    model.config.modality_dropout = 0.3
    meta["dropout_masks"] = masks

    # ---- eval forward, B=2, full shapes ---------------------------------------------------------------------------
    batch = ref_batch(batch_size=2, seed=1234)
    model.eval()
    torch.manual_seed(123)
    t0 = time.time()
    with torch.no_grad():
        y = model(batch)
        meta["eval_next_rand"] = torch.rand(1).item()  # three draws consumed even in eval (model.py:135-137)
        y_nopool = model(batch, pool_outputs=False)
        agg = model.aggregate_features(batch)
    print(f"eval forwards: {time.time() - t0:.1f}s")
    out["eval_y"] = y.numpy()
    out["eval_y_nopool_sub"] = y_nopool[:, ::50, :].numpy()
    out["eval_agg_sub"] = agg[:, ::37, ::101].numpy()

    # ---- validation step through the reference LightningModule -----------------------------------------------------
    metrics = torch.nn.ModuleDict({
        "val/pearson": MultidimPearsonCorrCoef(num_outputs=1000),
        "val/subj_pearson": GroupedMetric("MultidimPearsonCorrCoef", {"num_outputs": 1000}),
    })
    module = BrainModule(model=model, loss=torch.nn.MSELoss(), optim_config=None, metrics=metrics, max_epochs=1)
    module.eval()
    torch.manual_seed(123)
    with torch.no_grad():
        y_pred, y_true = module.validation_step(batch, 0)
    module.on_validation_epoch_end()
    meta["val_loss"] = float(module.logged["val/loss"])
    meta["val_pearson"] = float(metrics["val/pearson"].compute())
    meta["val_grouped"] = {k: float(v) for k, v in module.logged.items() if k.startswith("val/subj_pearson/")}
    assert y_pred.device.type == "cpu" and torch.equal(y_pred, y)

    # ---- train step: loss + per-parameter gradient norms (dropout active: seed chosen so one modality drops) --------
    module.train()
    seed = next(s for s in range(24) if len(masks[f"0.3:{s}"]) == 1)
    torch.manual_seed(seed)
    np.random.seed(seed)
    t0 = time.time()
    loss = module.training_step(batch, 0)
    loss.backward()
    print(f"train step: {time.time() - t0:.1f}s, seed={seed}, dropped={masks[f'0.3:{seed}']}")
    meta["train_seed"], meta["train_loss"], meta["train_dropped"] = seed, float(loss), masks[f"0.3:{seed}"]
    meta["grad_none"] = [k for k, v in model.named_parameters() if v.grad is None]
    names = [k for k, v in model.named_parameters() if v.grad is not None]
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array([dict(model.named_parameters())[k].grad.double().norm().item() for k in names])
    pd = dict(model.named_parameters())
    out["grad_predictor_bias"] = pd["predictor.bias"].grad.numpy()
    out["grad_time_pos_embed_sub"] = pd["time_pos_embed"].grad[0, :298:13, ::97].numpy()
    out["grad_l0_toq_sub"] = pd["encoder.layers.0.1.to_q.weight"].grad[::193, ::211].numpy()
    out["grad_l15_ff2_sub"] = pd["encoder.layers.15.1.ff.2.weight"].grad[::193, ::811].numpy()
    out["grad_residual_scale_l7"] = pd["encoder.layers.7.2.residual_scale"].grad.numpy()

    # ---- contrastive branch (defaults.py:102): second model, same seed, contrastive on ---------------------------
    del module, model
    torch.manual_seed(33)
    np.random.seed(33)
    cfg_c = FmriEncoderConfig(n_subjects=4, modality_dropout=0.0, contrastive_enabled=True)
    model_c = cfg_c.build(feature_dims=FEATURE_DIMS, n_outputs=1000, n_output_timesteps=100)
    model_c.eval()
    with torch.no_grad():
        closs = model_c.compute_contrastive_loss(batch)
    meta["contrastive_video"] = float(closs["video"])
    meta["contrastive_param_names"] = [k for k, _ in model_c.named_parameters() if "contrastive" in k]

    np.savez_compressed(os.path.join(OUT, "full_model.npz"), **out)
    with open(os.path.join(OUT, "full_model.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("full_model.npz / full_model.json written")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    small_ops()
    full_model()
