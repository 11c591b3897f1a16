"""Contrastive branch (reference algonauts2025/model.py:177-241, enabled by default in grids/defaults.py:102) on the
B200 kernels vs the reference's own InfoNCE value (tests/golden/small_ops.npz, full_model.json) and the CPU oracle."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import algonauts2025_b200  # noqa: E402,F401
from algonauts2025_b200.contrastive import info_nce  # noqa: E402
from algonauts2025_b200.model import FmriEncoder, FmriEncoderConfig  # noqa: E402
from algonauts2025_b200.pl_module import BrainModule  # noqa: E402
from algonauts2025_b200.segment import synthetic_batch  # noqa: E402
from oracle import tribe_oracle as O  # noqa: E402

SMALL_DIMS = {"text": (2, 96), "audio": (2, 40), "video": (1, 72)}
SMALL = dict(hidden=384, depth=2, heads=6)


def rel_l2(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_info_nce_matches_reference_value(golden_dir):
    g = np.load(os.path.join(golden_dir, "small_ops.npz"))
    q, k = torch.from_numpy(g["nce_q"]).cuda(), torch.from_numpy(g["nce_k"]).cuda()
    loss = info_nce(q, k, 0.07)
    assert abs(loss.item() - float(g["nce_loss"])) <= 2e-2 * abs(float(g["nce_loss"]))  # bf16 logits at 1/tau = 14.3


def test_info_nce_forward_backward_vs_oracle():
    torch.manual_seed(0)
    q = torch.randn(3, 50, 256, requires_grad=True)
    k = (0.5 * q.detach() + torch.randn(3, 50, 256)).requires_grad_(True)
    ref = O.OracleFmriEncoder.info_nce(q, k, 0.07)
    ref.backward()
    qc, kc = q.detach().cuda().requires_grad_(True), k.detach().cuda().requires_grad_(True)
    loss = info_nce(qc, kc, 0.07)
    (2.0 * loss).backward()
    assert abs(loss.item() - ref.item()) <= 2e-2 * abs(ref.item())
    assert rel_l2(qc.grad / 2.0, q.grad) <= 5e-2
    assert rel_l2(kc.grad / 2.0, k.grad) <= 5e-2


def test_contrastive_train_step_matches_oracle():
    cfg_kw = dict(n_subjects=3, contrastive_enabled=True, modality_dropout=0.0)
    torch.manual_seed(5)
    model = FmriEncoder(SMALL_DIMS, 200, 25, FmriEncoderConfig(**cfg_kw), **SMALL)
    oracle = O.OracleFmriEncoder(SMALL_DIMS, 200, 25, O.OracleConfig(**cfg_kw), **SMALL)
    oracle.load_reference_state_dict({k: v.detach().cpu().clone() for k, v in model.state_dict().items()})
    batch = synthetic_batch(batch_size=3, t=74, t_out=25, n_outputs=200, n_subjects=3, seed=7,
                            dims=(("text", 2, 96), ("audio", 2, 40), ("video", 1, 72)))
    module = BrainModule(model=model, loss=torch.nn.MSELoss(), optim_config=None, metrics={}, max_epochs=1)
    module.train(), oracle.train()
    torch.manual_seed(3), np.random.seed(3)
    loss = module.training_step(batch, 0)
    loss.backward()
    torch.manual_seed(3), np.random.seed(3)
    ref_loss, _, _, ref_c = O.run_step(oracle, O.SegmentData(batch.data, batch.segments))
    ref_loss.backward()
    assert abs(float(module.logged["train/contrastive/video"]) - ref_c["video"].item()) <= 2e-2 * ref_c["video"].item()
    assert abs(loss.item() - ref_loss.item()) <= 1e-2 * ref_loss.item()
    ref_grads = {k.replace("predictor_weights", "predictor.weights").replace("predictor_bias", "predictor.bias"): v.grad
                 for k, v in oracle.named_parameters()}
    for name, p in model.named_parameters():
        assert p.grad is not None, name
        # encoder weights receive the sum of two backward passes (prediction + contrastive latents)
        assert rel_l2(p.grad, ref_grads[name]) <= 5e-2, (name, rel_l2(p.grad, ref_grads[name]))


def test_full_model_contrastive_value_matches_reference(golden_dir):
    """Seed 33, contrastive head enabled, eval mode: the reference's own compute_contrastive_loss()['video']."""
    meta = json.load(open(os.path.join(golden_dir, "full_model.json")))
    torch.manual_seed(33)
    np.random.seed(33)
    cfg = FmriEncoderConfig(n_subjects=4, modality_dropout=0.0, contrastive_enabled=True)
    model = cfg.build(feature_dims={"text": (2, 3072), "audio": (2, 1024), "video": (2, 1408)}, n_outputs=1000, n_output_timesteps=100)
    assert [n for n, _ in model.named_parameters() if "contrastive" in n] == meta["contrastive_param_names"]
    batch = synthetic_batch(batch_size=2, seed=1234)
    model.eval()
    with torch.no_grad():
        closs = model.compute_contrastive_loss(batch)
    assert abs(closs["video"].item() - meta["contrastive_video"]) <= 2e-2 * meta["contrastive_video"]
