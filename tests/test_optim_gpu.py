"""TribeAdam (fused Adam + bf16 shadow kernel) vs stock torch.optim.Adam under OneCycleLR, including parameters that
skip steps because their gradient is None (dropped modality) and the adopted-instance plumbing."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import algonauts2025_b200  # noqa: E402,F401
from algonauts2025_b200.model import FmriEncoder, FmriEncoderConfig  # noqa: E402
from algonauts2025_b200.optim import TribeAdam  # noqa: E402

DIMS = {"text": (2, 96), "audio": (2, 40), "video": (1, 72)}


def test_fused_adam_matches_torch_adam_with_onecycle_and_skipped_params():
    torch.manual_seed(0)
    model = FmriEncoder(DIMS, 50, 10, FmriEncoderConfig(n_subjects=2), hidden=384, depth=1, heads=6)
    eng = model._engine
    eng._check_flat()
    eng.flat.ensure_grad()
    ref = {n: p.detach().clone().requires_grad_(True) for n, p in model.named_parameters()}
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-3, pct_start=0.3, total_steps=8)
    TribeAdam.adopt(opt, model)
    assert type(opt) is TribeAdam and isinstance(opt, torch.optim.Adam)
    ropt = torch.optim.Adam(list(ref.values()), lr=1e-3)
    rsched = torch.optim.lr_scheduler.OneCycleLR(ropt, max_lr=1e-3, pct_start=0.3, total_steps=8)
    g = torch.Generator(device="cuda").manual_seed(1)
    for step in range(6):
        skip = {"projectors.video.weight", "projectors.video.bias"} if step in (1, 4) else set()
        for n, p in model.named_parameters():
            if n in skip:
                p.grad, ref[n].grad = None, None
                continue
            grad = torch.randn(p.shape, device="cuda", generator=g) * (0.1 + step)
            gv = eng.flat.gview(n)
            gv.copy_(grad)
            p.grad = gv if step % 2 == 0 else grad.clone()  # flat view and foreign gradient tensors both work
            ref[n].grad = grad.clone()
        opt.step(), sched.step()
        ropt.step(), rsched.step()
        assert opt.param_groups[0]["lr"] == ropt.param_groups[0]["lr"]
    for n, p in model.named_parameters():
        torch.testing.assert_close(p.detach(), ref[n].detach(), rtol=2e-5, atol=1e-7, msg=n)
        st, rst = opt.state[p], ropt.state[ref[n]]
        assert float(st["step"]) == float(rst["step"]), n
        torch.testing.assert_close(st["exp_avg"], rst["exp_avg"], rtol=2e-5, atol=5e-6)  # |grad| up to ~5: fp32 rounding of the lerp
        torch.testing.assert_close(st["exp_avg_sq"], rst["exp_avg_sq"], rtol=2e-5, atol=1e-5)
        # the bf16 shadow was written by the same kernel
        assert torch.equal(eng.flat.view16(n), p.detach().to(torch.bfloat16)), n
    # the engine does not re-cast after a fused step ...
    casts = []
    orig = algonauts2025_b200.ops.cast_f32_bf16
    algonauts2025_b200.ops.cast_f32_bf16 = lambda *a, **k: casts.append(1) or orig(*a, **k)
    try:
        eng.flat.refresh_bf16()
        assert not casts
        with torch.no_grad():
            model.predictor.bias.add_(1.0)  # ... but does after any other in-place change
        eng.flat.refresh_bf16()
        assert casts
    finally:
        algonauts2025_b200.ops.cast_f32_bf16 = orig
    sd = opt.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}
