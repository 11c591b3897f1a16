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


@pytest.mark.parametrize("m,n,k,accumulate", [(1024, 512, 320, False), (1096, 520, 256, False), (1280, 768, 192, True), (384, 200, 128, False)])
def test_adam_in_the_wgrad_gemm_epilogue_is_bit_identical_to_gemm_then_adam(m, n, k, accumulate):
    """TribeGemm.adam_* (include/tribe_b200.h): the optimizer step applied to the finished gradient tile inside the GEMM
    epilogue — 2-CTA kernel through the coalescing transpose buffer for m >= 1024 (full and ragged row / column tiles),
    1-CTA kernel row-per-thread otherwise, with and without in-place gradient accumulation — against the same GEMM
    followed by the stand-alone Adam kernel: parameters, both moments and the bf16 shadow must agree bit for bit."""
    import ctypes

    from algonauts2025_b200 import _lib, ops

    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(m + n)
    a = (torch.randn(k, m, device="cuda", generator=g) * 0.5).to(torch.bfloat16)  # wgrad form: both operands MN-major
    b = (torch.randn(k, n, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    p0 = torch.randn(m, n, device="cuda", generator=g)
    m0 = torch.randn(m, n, device="cuda", generator=g) * 0.1
    v0 = torch.rand(m, n, device="cuda", generator=g) * 0.01
    g_prev = torch.randn(m, n, device="cuda", generator=g)
    hyper = torch.zeros(8, device="cuda")
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.tribe_adam_hyper(ctypes.c_void_p(hyper.data_ptr()), 1e-3, 0.9, 0.999, 1e-8, 0.01, 3, stream), "tribe_adam_hyper")

    def grad_gemm(out, **kw):
        if accumulate:
            ops.gemm(ops.mnmajor(a), ops.mnmajor(b), out, m, n, k, ldd=n, epilogue=ops.EPI_RESIDUAL, res=out, ld_res=n, res_batched=True, **kw)
        else:
            ops.gemm(ops.mnmajor(a), ops.mnmajor(b), out, m, n, k, ldd=n, **kw)

    # reference: gradient GEMM, then the flat Adam kernel
    gr = g_prev.clone()
    pr, mr, vr, sr = p0.clone(), m0.clone(), v0.clone(), torch.zeros(m, n, device="cuda", dtype=torch.bfloat16)
    grad_gemm(gr)
    _lib.check(lib.tribe_adam_step_dev(*(ctypes.c_void_p(t.data_ptr()) for t in (pr, gr, mr, vr, sr)), m * n, ctypes.c_void_p(hyper.data_ptr()), 0, stream),
               "tribe_adam_step_dev")
    for keep in (False, True):
        gf = g_prev.clone()
        pf, mf, vf, sf = p0.clone(), m0.clone(), v0.clone(), torch.zeros(m, n, device="cuda", dtype=torch.bfloat16)
        grad_gemm(gf, adam=(pf.data_ptr(), mf.data_ptr(), vf.data_ptr(), sf.data_ptr(), hyper.data_ptr(), keep))
        torch.cuda.synchronize()
        assert torch.equal(pf, pr) and torch.equal(mf, mr) and torch.equal(vf, vr) and torch.equal(sf, sr), (m, n, k, keep)
        assert torch.equal(gf, gr if keep else g_prev)  # the gradient is written only on request
    with pytest.raises(algonauts2025_b200.TribeError):  # needs a row-major fp32 output
        ops.gemm(ops.mnmajor(a), ops.mnmajor(b), torch.empty(m, n, device="cuda", dtype=torch.bfloat16), m, n, k, ldd=n,
                 adam=(pf.data_ptr(), mf.data_ptr(), vf.data_ptr(), sf.data_ptr(), hyper.data_ptr(), False))
