"""Parity of every C-ABI kernel (called through algonauts2025_b200.ops -> ctypes -> libtribe_b200.so) against the CPU
oracle / plain torch fp32 references.  Tolerances: bit-exact for index/gather work, 1e-2 relative for bf16 GEMM paths
(BASELINE.json north_star), 1e-5 for fp32 bandwidth kernels, 1e-3 abs on Pearson r (observed ~1e-6)."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import algonauts2025_b200  # noqa: E402,F401
from algonauts2025_b200 import ops  # noqa: E402
from oracle import tribe_oracle as O  # noqa: E402

DEV = "cuda"


def bf(x):
    return x.to(torch.bfloat16)


def assert_close_bf16(out, ref, rtol=1e-2, atol=None):
    ref = ref.float()
    atol = atol if atol is not None else 1e-2 * float(ref.abs().max())
    err = (out.float() - ref).abs()
    assert torch.isfinite(out.float()).all()
    assert bool((err <= atol + rtol * ref.abs()).all()), f"max err {float(err.max())} (ref max {float(ref.abs().max())})"


# ------------------------------------------------------------------------------------------------------------ GEMM
def _mn_operand(x):
    """(MN, K) matrix stored transposed as (K, MN) with a 16-byte aligned (padded) row stride: an MN-major operand."""
    mn, k = x.shape
    pad = (mn + 7) // 8 * 8
    store = torch.zeros(k, pad, device=x.device, dtype=x.dtype)
    store[:, :mn] = x.t()
    return ops.Operand(store, inner=mn, rows=k, row_stride=pad, mn_major=True)


@pytest.mark.parametrize("a_mn", [False, True])
@pytest.mark.parametrize("b_mn", [False, True])
@pytest.mark.parametrize("bn", [0, 128, 160, 192, 256])
def test_gemm_majors_and_tiles(a_mn, b_mn, bn):
    if bn == 160 and b_mn:
        pytest.skip("block_n=160 needs K-major B")
    torch.manual_seed(1)
    m, n, k = 333, 520, 200  # ragged in every dimension
    A, B = bf(torch.randn(m, k, device=DEV)), bf(torch.randn(n, k, device=DEV))
    a_op = _mn_operand(A) if a_mn else ops.kmajor(A)
    b_op = _mn_operand(B) if b_mn else ops.kmajor(B)
    out = torch.full((m, n), float("nan"), device=DEV)
    ops.gemm(a_op, b_op, out, m, n, k, ldd=n, block_n=bn)
    assert_close_bf16(out, A.float() @ B.float().t())


def test_gemm_large_persistent_multi_tile():
    """More tiles than SMs, K long enough to wrap the smem ring many times, bf16 output."""
    torch.manual_seed(2)
    m, n, k = 4768, 3072, 1024
    A, B = bf(torch.randn(m, k, device=DEV)), bf(torch.randn(n, k, device=DEV) / math.sqrt(k))
    out = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    ops.linear(A, B, out)
    assert_close_bf16(out, A.float() @ B.float().t())


def test_gemm_epilogues_bias_gelu_residual():
    torch.manual_seed(3)
    m, n, k, T = 300, 384, 256, 50
    A, W = bf(torch.randn(m, k, device=DEV)), bf(torch.randn(n, k, device=DEV) / 16)
    bias = torch.randn(n, device=DEV)
    acc = A.float() @ W.float().t() + bias
    # GELU: aux_out gets the pre-activation
    out, aux = torch.empty(m, n, device=DEV, dtype=torch.bfloat16), torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    ops.linear(A, W, out, bias=bias, epilogue=ops.EPI_GELU, aux_out=aux, ld_aux=n)
    assert_close_bf16(aux, acc)
    assert_close_bf16(out, torch.nn.functional.gelu(acc))
    # GELU backward: D = acc * gelu'(aux)
    hpre = bf(torch.randn(m, n, device=DEV))
    out_b = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    ops.linear(A, W, out_b, epilogue=ops.EPI_GELU_BWD, aux_in=hpre, ld_aux=n)
    h = hpre.float().requires_grad_(True)
    torch.nn.functional.gelu(h).sum().backward()
    assert_close_bf16(out_b, (A.float() @ W.float().t()) * h.grad)
    # residual with per-column scale and periodic residual rows (positional-embedding add)
    res, rs = torch.randn(T, n, device=DEV), torch.rand(n, device=DEV) + 0.5
    out_r = torch.empty(m, n, device=DEV)
    ops.linear(A, W, out_r, bias=bias, epilogue=ops.EPI_RESIDUAL, res=res, ld_res=n, res_row_mod=T, rscale=rs)
    rows = torch.arange(m, device=DEV) % T
    assert_close_bf16(out_r, acc + res[rows] * rs)
    # strided output (column block of a wider buffer), as used for the concatenated projector outputs
    wide = torch.zeros(m, 3 * n, device=DEV)
    ops.linear(A, W, wide, ldd=3 * n, d_off=n, bias=bias)
    assert_close_bf16(wide[:, n:2 * n], acc)
    assert float(wide[:, :n].abs().max()) == 0.0 and float(wide[:, 2 * n:].abs().max()) == 0.0


@pytest.mark.parametrize("m,n", [(1280, 768), (1100, 520)])
def test_gemm_2cta_staged_epilogues(m, n):
    """The 2-CTA kernel's round-2 epilogues — bf16 outputs through the shared-memory tile + TMA store (plain, GELU with and
    without its pre-activation side output, GELU'), fp32 outputs through the transposed coalesced path (bias, periodic
    residual rows x column scale, in-place accumulation) — on full tiles and on ragged rows / a ragged last chunk."""
    torch.manual_seed(m + n)
    k, T = 320, 100
    A, W = bf(torch.randn(m, k, device=DEV)), bf(torch.randn(n, k, device=DEV) / 16)
    bias = torch.randn(n, device=DEV)
    acc = A.float() @ W.float().t() + bias
    out, aux = torch.full((m, n), float("nan"), device=DEV, dtype=torch.bfloat16), torch.full((m, n), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.linear(A, W, out, bias=bias, epilogue=ops.EPI_GELU, aux_out=aux, ld_aux=n)
    assert_close_bf16(aux, acc)
    assert_close_bf16(out, torch.nn.functional.gelu(acc))
    out_inf = torch.full((m, n), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.linear(A, W, out_inf, bias=bias, epilogue=ops.EPI_GELU)  # inference: no pre-activation output
    assert torch.equal(out_inf, out)
    hpre = bf(torch.randn(m, n, device=DEV))
    out_b = torch.full((m, n), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.linear(A, W, out_b, epilogue=ops.EPI_GELU_BWD, aux_in=hpre, ld_aux=n)
    h = hpre.float().requires_grad_(True)
    torch.nn.functional.gelu(h).sum().backward()
    assert_close_bf16(out_b, (A.float() @ W.float().t()) * h.grad)
    res, rs = torch.randn(T, n, device=DEV), torch.rand(n, device=DEV) + 0.5
    out_r = torch.full((m, n), float("nan"), device=DEV)
    ops.linear(A, W, out_r, bias=bias, epilogue=ops.EPI_RESIDUAL, res=res, ld_res=n, res_row_mod=T, rscale=rs)
    assert_close_bf16(out_r, acc + res[torch.arange(m, device=DEV) % T] * rs)
    g0 = torch.randn(m, n, device=DEV)
    g = g0.clone()
    ops.linear(A, W, g, epilogue=ops.EPI_RESIDUAL, res=g, ld_res=n, res_batched=True)  # D += acc (second-pass gradient accumulation)
    assert_close_bf16(g - g0, A.float() @ W.float().t())
    wide = torch.zeros(m, 2 * n + 8, device=DEV, dtype=torch.bfloat16)
    ops.linear(A, W, wide, ldd=2 * n + 8, d_off=8, bias=bias)  # column block of a wider bf16 buffer (TMA store with an offset base)
    assert_close_bf16(wide[:, 8:8 + n], acc)
    assert float(wide[:, :8].abs().max()) == 0.0 and float(wide[:, 8 + n:].abs().max()) == 0.0


def _rope_table(T, rot_dim):
    inv = 1.0 / (10000 ** (torch.arange(0, rot_dim, 2).float() / rot_dim))
    ang = torch.arange(T).float()[:, None] * inv[None, :]
    return torch.stack((ang.cos(), ang.sin()), dim=-1).contiguous()  # (T, rot_dim/2, 2)


def test_gemm_rope_epilogue_matches_oracle():
    from oracle.xt_encoder import RotaryEmbedding, apply_rotary

    torch.manual_seed(4)
    Bsz, T, heads, dh, rot = 2, 37, 2, 64, 32
    dim = heads * dh
    m, k = Bsz * T, 128
    X, W = bf(torch.randn(m, k, device=DEV)), bf(torch.randn(3 * dim, k, device=DEV) / 11)
    table = _rope_table(T, rot).to(DEV)
    out = torch.empty(m, 3 * dim, device=DEV, dtype=torch.bfloat16)
    ops.linear(X, W, out, epilogue=ops.EPI_ROPE, rope=table, rope_t=T, rope_dim=rot, head_dim=dh, rope_cols=2 * dim)
    qkv = (X.float() @ W.float().t()).cpu().view(Bsz, T, 3, heads, dh)
    freqs = RotaryEmbedding(rot)(T)
    q = apply_rotary(qkv[:, :, 0].transpose(1, 2), freqs).transpose(1, 2)
    kk = apply_rotary(qkv[:, :, 1].transpose(1, 2), freqs).transpose(1, 2)
    ref = torch.stack((q, kk, qkv[:, :, 2]), dim=2).reshape(m, 3 * dim)
    assert_close_bf16(out.cpu(), ref)
    # inverse rotation (backward): R^T R = I
    G = bf(torch.randn(m, k, device=DEV))
    eye = bf(torch.eye(dim, device=DEV))  # acc = G[:, :dim] via identity weights
    g_rot = torch.empty(m, dim, device=DEV)
    Gd = bf(torch.randn(m, dim, device=DEV))
    ops.linear(Gd, eye, g_rot, epilogue=ops.EPI_ROPE, rope=table, rope_t=T, rope_dim=rot, head_dim=dh, rope_cols=dim)
    back = torch.empty(m, dim, device=DEV)
    ops.linear(bf(g_rot), eye, back, epilogue=ops.EPI_ROPE, rope=table, rope_t=T, rope_dim=rot, head_dim=dh, rope_cols=dim, rope_sign=-1.0)
    assert_close_bf16(back, Gd.float(), rtol=2e-2)


def test_gemm_batched_attention_shapes():
    """S = Q K^T and O = P V per (b, h) straight out of a packed [tokens, 3*H*dh] buffer; dV = P^T dO (both MN-major)."""
    torch.manual_seed(5)
    Bsz, T, H, dh = 2, 70, 3, 64
    dim = H * dh
    qkv = bf(torch.randn(Bsz * T, 3 * dim, device=DEV))
    Tp = 80  # padded key length (multiple of 8 -> 16-byte rows)
    S = torch.full((Bsz * H, T, Tp), float("nan"), device=DEV)
    q_op = ops.Operand(qkv, inner=3 * dim, rows=T, row_stride=3 * dim, batch=Bsz, batch_stride=T * 3 * dim, zin_stride=dh, zdiv=H)
    k_op = ops.Operand(qkv, inner=3 * dim, rows=T, row_stride=3 * dim, batch=Bsz, batch_stride=T * 3 * dim, inner_off=dim, zin_stride=dh, zdiv=H)
    ops.gemm(q_op, k_op, S, T, T, dh, ldd=Tp, batch=Bsz * H, z_inner=H, d_zo=H * T * Tp, d_zi=T * Tp, alpha=dh ** -0.5)
    q4 = qkv.float().view(Bsz, T, 3, H, dh)
    ref_S = torch.einsum("bihd,bjhd->bhij", q4[:, :, 0], q4[:, :, 1]) * dh ** -0.5
    assert_close_bf16(S.view(Bsz, H, T, Tp)[..., :T], ref_S)
    # softmax -> P (bf16, zero padded), then O = P V with V as an MN-major operand inside qkv
    P = torch.empty(Bsz * H, T, Tp, device=DEV, dtype=torch.bfloat16)
    ops.softmax_fwd(S, P, T)
    ref_P = ref_S.softmax(-1)
    assert_close_bf16(P.view(Bsz, H, T, Tp)[..., :T], ref_P, atol=2e-3)
    assert float(P.view(Bsz, H, T, Tp)[..., T:].float().abs().max()) == 0.0
    attn = torch.full((Bsz * T, dim), float("nan"), device=DEV, dtype=torch.bfloat16)
    p_op = ops.Operand(P, inner=Tp, rows=T, row_stride=Tp, batch=Bsz * H, batch_stride=T * Tp)
    v_op = ops.Operand(qkv, inner=3 * dim, rows=T, row_stride=3 * dim, batch=Bsz, batch_stride=T * 3 * dim, mn_major=True,
                       inner_off=2 * dim, zin_stride=dh, zdiv=H)
    ops.gemm(p_op, v_op, attn, T, dh, Tp, ldd=dim, batch=Bsz * H, z_inner=H, d_zo=T * dim, d_zi=dh)
    ref_O = torch.einsum("bhij,bjhd->bihd", P.float().view(Bsz, H, T, Tp)[..., :T], q4[:, :, 2]).reshape(Bsz * T, dim)
    assert_close_bf16(attn, ref_O)
    # dV = P^T dO : A = P^T (MN-major, M = keys), B = dO (MN-major, N = head dims), K = queries
    dO = bf(torch.randn(Bsz * T, dim, device=DEV))
    dqkv = torch.zeros(Bsz * T, 3 * dim, device=DEV, dtype=torch.bfloat16)
    pt_op = ops.Operand(P, inner=Tp, rows=T, row_stride=Tp, batch=Bsz * H, batch_stride=T * Tp, mn_major=True)
    do_op = ops.Operand(dO, inner=dim, rows=T, row_stride=dim, batch=Bsz, batch_stride=T * dim, mn_major=True, zin_stride=dh, zdiv=H)
    ops.gemm(pt_op, do_op, dqkv, T, dh, T, ldd=3 * dim, batch=Bsz * H, z_inner=H, d_zo=T * 3 * dim, d_zi=dh, d_off=2 * dim)
    ref_dV = torch.einsum("bhij,bihd->bjhd", P.float().view(Bsz, H, T, Tp)[..., :T], dO.float().view(Bsz, T, H, dh)).reshape(Bsz * T, dim)
    assert_close_bf16(dqkv[:, 2 * dim:], ref_dV)
    assert float(dqkv[:, :2 * dim].float().abs().max()) == 0.0


def test_gemm_subject_gather_readout_and_grouped_wgrad():
    """SubjectLayers (common.py:45-67): per-sample weights selected by subject id inside the TMA coordinates."""
    torch.manual_seed(6)
    Bsz, Tq, C, Oc, S = 5, 20, 128, 200, 4
    subj = torch.tensor([3, 0, 3, 1, 0], device=DEV)
    W = torch.randn(S, C, Oc, device=DEV) / math.sqrt(C)
    bias = torch.randn(S, Oc, device=DEV)
    x = bf(torch.randn(Bsz, Tq, C, device=DEV))
    Wb = bf(W)
    out = torch.full((Bsz, Oc, Tq), float("nan"), device=DEV)
    x_op = ops.Operand(x, inner=C, rows=Tq, row_stride=C, batch=Bsz, batch_stride=Tq * C)
    w_op = ops.Operand(Wb, inner=Oc, rows=C, row_stride=Oc, batch=S, batch_stride=C * Oc, mn_major=True, gather=subj)
    ops.gemm(x_op, w_op, out, Tq, Oc, C, ldd=Tq, batch=Bsz, d_zo=Oc * Tq, transposed=True, bias=bias, bias_gathered=True, bias_z_stride=Oc)
    ref = O.subject_layers(x.float().cpu().transpose(1, 2), subj.cpu()[:, None], Wb.float().cpu(), bias.cpu())
    assert_close_bf16(out.cpu(), ref)
    # grouped wgrad: dW[s] = sum_{b: subj[b]==s} x[b]^T dy[b]; subject 2 never appears -> exact zeros
    dy = bf(torch.randn(Bsz, Tq, Oc, device=DEV))
    dW = torch.full((S, C, Oc), float("nan"), device=DEV)
    xa = ops.Operand(x, inner=C, rows=Tq, row_stride=C, batch=Bsz, batch_stride=Tq * C, mn_major=True)
    dyb = ops.Operand(dy, inner=Oc, rows=Tq, row_stride=Oc, batch=Bsz, batch_stride=Tq * Oc, mn_major=True)
    ops.gemm(xa, dyb, dW, C, Oc, Tq, ldd=Oc, batch=S, d_zo=C * Oc, kgroup=subj)
    ref_dW = torch.zeros(S, C, Oc)
    for b in range(Bsz):
        ref_dW[subj[b].item()] += x[b].float().cpu().t() @ dy[b].float().cpu()
    assert_close_bf16(dW.cpu(), ref_dW)
    assert float(dW[2].abs().max()) == 0.0
    # bias gradient
    db = torch.zeros(S, Oc, device=DEV)
    ops.subject_bias_grad(dy, subj, db, Bsz, Tq, Oc, S)
    ref_db = torch.zeros(S, Oc)
    ref_db.index_add_(0, subj.cpu(), dy.float().cpu().sum(1))
    torch.testing.assert_close(db.cpu(), ref_db, rtol=1e-4, atol=1e-4)
    # subject range check flag (assert at common.py:53-55)
    flag = torch.zeros(1, device=DEV, dtype=torch.int32)
    ops.check_subjects(subj, S, flag)
    assert int(flag.item()) == 0
    ops.check_subjects(subj, 3, flag)
    assert int(flag.item()) == 1


# ------------------------------------------------------------------------------------------------- bandwidth kernels
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64, torch.bfloat16])
@pytest.mark.parametrize("mean", [False, True])
def test_ingest_features(dtype, mean):
    torch.manual_seed(7)
    Bsz, L, D, T = 3, 2, 100, 77
    x = torch.randn(Bsz, L, D, T, device=DEV).to(dtype)
    width = (D if mean else L * D)
    out = torch.zeros(Bsz * T, width + 24, device=DEV, dtype=torch.bfloat16)
    ops.ingest_features(x, out, 16, mean)
    xf = x.float()
    ref = (xf.mean(1) if mean else xf.reshape(Bsz, L * D, T)).transpose(1, 2).reshape(Bsz * T, width)
    assert torch.equal(out[:, 16:16 + width], ref.to(torch.bfloat16))  # a cast + a gather: bit-exact
    assert float(out[:, :16].float().abs().max()) == 0.0 and float(out[:, 16 + width:].float().abs().max()) == 0.0


def test_ingest_features_3d_input():
    x = torch.randn(2, 40, 33, device=DEV)
    out = torch.zeros(2 * 33, 40, device=DEV, dtype=torch.bfloat16)
    ops.ingest_features(x, out, 0, False)
    assert torch.equal(out, x.transpose(1, 2).reshape(66, 40).to(torch.bfloat16))


def test_scalenorm_forward_and_sublayer_backward():
    torch.manual_seed(8)
    rows, dim = 130, 3072
    x = torch.randn(rows, dim, device=DEV) * 3
    g = torch.tensor([1.3], device=DEV)
    y, rn = torch.empty(rows, dim, device=DEV, dtype=torch.bfloat16), torch.empty(rows, device=DEV)
    ops.scalenorm_fwd(x, g, y, rn)
    ref = torch.nn.functional.normalize(x, dim=-1) * dim ** 0.5 * g
    assert_close_bf16(y, ref)
    torch.testing.assert_close(rn, 1 / x.norm(dim=-1), rtol=1e-5, atol=0)
    # backward tail vs autograd of  out = f(norm(x)) + x * rs  with upstream grads d_xn (wrt norm output), dy_out
    rs = torch.rand(dim, device=DEV) + 0.5
    d_xn, dy_out = bf(torch.randn(rows, dim, device=DEV)), torch.randn(rows, dim, device=DEV)
    xr, gr, rsr = x.clone().requires_grad_(True), g.clone().requires_grad_(True), rs.clone().requires_grad_(True)
    yn = torch.nn.functional.normalize(xr, dim=-1) * dim ** 0.5 * gr
    ((yn * d_xn.float()).sum() + (xr * rsr * dy_out).sum()).backward()
    dx, dxb = torch.empty(rows, dim, device=DEV), torch.empty(rows, dim, device=DEV, dtype=torch.bfloat16)
    d_rs, d_g = torch.zeros(dim, device=DEV), torch.zeros(1, device=DEV)
    ops.sublayer_bwd(dy_out, d_xn, x, rn, g, rs, dx, dxb, d_rs, d_g)
    torch.testing.assert_close(dx, xr.grad, rtol=1e-4, atol=1e-4)
    assert_close_bf16(dxb, xr.grad)
    torch.testing.assert_close(d_rs, rsr.grad, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(d_g, gr.grad, rtol=1e-3, atol=1e-2)
    # final-norm form: no residual path
    dx2 = torch.empty(rows, dim, device=DEV)
    d_g2 = torch.zeros(1, device=DEV)
    ops.sublayer_bwd(None, d_xn, x, rn, g, None, dx2, None, None, d_g2)
    xr2 = x.clone().requires_grad_(True)
    (torch.nn.functional.normalize(xr2, dim=-1) * dim ** 0.5 * g * d_xn.float()).sum().backward()
    torch.testing.assert_close(dx2, xr2.grad, rtol=1e-4, atol=1e-4)


def test_softmax_backward():
    torch.manual_seed(9)
    rows, n, ld = 100, 298, 304
    s = torch.randn(rows, ld, device=DEV) * 2
    p = torch.empty(rows, ld, device=DEV, dtype=torch.bfloat16)
    ops.softmax_fwd(s, p, n)
    torch.testing.assert_close(p[:, :n].float(), s[:, :n].softmax(-1), rtol=1e-2, atol=1e-4)
    dp = torch.randn(rows, ld, device=DEV)
    ds = torch.full((rows, ld), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.softmax_bwd(p, dp, ds, 0.25, n)
    pf = p[:, :n].float()
    ref = pf * (dp[:, :n] - (pf * dp[:, :n]).sum(-1, keepdim=True)) * 0.25
    assert_close_bf16(ds[:, :n], ref)
    assert float(ds[:, n:].float().abs().max()) == 0.0


def test_colsum_cast_axpby():
    torch.manual_seed(10)
    x, y = torch.randn(1000, 300, device=DEV), torch.randn(1000, 300, device=DEV)
    out = torch.full((300,), float("nan"), device=DEV)
    ops.colsum(x, out)
    torch.testing.assert_close(out, x.sum(0), rtol=1e-4, atol=1e-3)
    ops.colsum(x, out, y=y, accumulate=True)
    torch.testing.assert_close(out, x.sum(0) + (x * y).sum(0), rtol=1e-4, atol=1e-3)
    xb = bf(x)
    ops.colsum(xb, out)
    torch.testing.assert_close(out, xb.float().sum(0), rtol=1e-4, atol=1e-3)
    ops.colsum(x, out, y=xb)
    torch.testing.assert_close(out, (x * xb.float()).sum(0), rtol=1e-4, atol=1e-3)
    src = torch.randn(100003 * 8, device=DEV)[: 100003 * 8 - 5]
    src = src[: (src.numel() // 8) * 8 + 3].clone()
    dst = torch.empty(src.numel(), device=DEV, dtype=torch.bfloat16)
    ops.cast_f32_bf16(src, dst)
    assert torch.equal(dst, src.to(torch.bfloat16))
    acc = torch.ones(777, device=DEV)
    ops.axpby(torch.arange(777, device=DEV, dtype=torch.float32), acc, 0.5, accumulate=True)
    torch.testing.assert_close(acc, 1 + 0.5 * torch.arange(777, device=DEV, dtype=torch.float32))


@pytest.mark.parametrize("t_in,t_out", [(298, 100), (300, 100), (97, 100), (250, 7)])
def test_adaptive_pool_golden_and_backward(golden_dir, t_in, t_out):
    g = np.load(os.path.join(golden_dir, "small_ops.npz"))
    x = torch.from_numpy(g[f"pool_{t_in}_{t_out}_x"]).to(DEV)
    y = ops.adaptive_avg_pool_fwd(x, t_out)
    np.testing.assert_allclose(y.cpu().numpy(), g[f"pool_{t_in}_{t_out}_y"], rtol=1e-6, atol=1e-6)  # reference's own output
    big = torch.randn(16, 1000, t_in, device=DEV, requires_grad=True)
    yb = torch.nn.AdaptiveAvgPool1d(t_out)(big)
    torch.testing.assert_close(ops.adaptive_avg_pool_fwd(big.detach(), t_out), yb.detach(), rtol=1e-6, atol=1e-6)
    dy = torch.randn_like(yb)
    yb.backward(dy)
    torch.testing.assert_close(ops.adaptive_avg_pool_bwd(dy, t_in), big.grad, rtol=1e-6, atol=1e-6)


def test_token_pool_matches_channel_pool():
    torch.manual_seed(11)
    Bsz, T, Tq, C = 3, 298, 100, 256
    x = bf(torch.randn(Bsz, T, C, device=DEV))
    y = torch.empty(Bsz, Tq, C, device=DEV, dtype=torch.bfloat16)
    ops.token_pool_fwd(x, y, Bsz, T, Tq, C)
    ref = O.adaptive_avg_pool1d(x.float().cpu().transpose(1, 2), Tq).transpose(1, 2)
    assert_close_bf16(y.cpu(), ref, atol=1e-2)
    dy = torch.randn(Bsz, Tq, C, device=DEV)
    dx = torch.empty(Bsz, T, C, device=DEV)
    ops.token_pool_bwd(dy, dx, Bsz, T, Tq, C)
    xr = x.float().transpose(1, 2).clone().requires_grad_(True)
    torch.nn.AdaptiveAvgPool1d(Tq)(xr).backward(dy.transpose(1, 2))
    torch.testing.assert_close(dx, xr.grad.transpose(1, 2), rtol=1e-5, atol=1e-6)
    dxb = torch.empty(Bsz, T, C, device=DEV)
    ops.token_pool_bwd(bf(dy), dxb, Bsz, T, Tq, C)
    assert_close_bf16(dxb, xr.grad.transpose(1, 2), atol=1e-2)


def test_transpose_cast_bot():
    x = torch.randn(3, 70, 45, device=DEV)
    y = torch.empty(3, 45, 70, device=DEV, dtype=torch.bfloat16)
    ops.transpose_cast_bot(x, y)
    assert torch.equal(y, x.transpose(1, 2).to(torch.bfloat16))


# ------------------------------------------------------------------------------------------------- loss / evaluation
def test_mse_forward_backward():
    torch.manual_seed(12)
    pred = torch.randn(16, 1000, 100, device=DEV, requires_grad=True)
    target = torch.randn(16, 1000, 100, device=DEV)
    ref = torch.nn.functional.mse_loss(pred, target)
    ref.backward()
    loss, grad = ops.mse_fwd_bwd(pred.detach(), target)
    torch.testing.assert_close(loss[0], ref.detach(), rtol=1e-6, atol=0)
    torch.testing.assert_close(grad, pred.grad, rtol=1e-6, atol=1e-12)
    odd = torch.randn(1237, device=DEV)
    l2, _ = ops.mse_fwd_bwd(odd, torch.zeros_like(odd), want_grad=False)
    torch.testing.assert_close(l2[0], (odd ** 2).mean(), rtol=1e-6, atol=0)


def test_pearson_small_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "small_ops.npz"))
    p, t = torch.from_numpy(g["loss_pred"]).to(DEV), torch.from_numpy(g["loss_true"]).to(DEV)
    stats = torch.zeros(1, 6, 17, device=DEV, dtype=torch.float64)
    for sl in (slice(0, 10), slice(10, 45), slice(45, 64)):  # streamed like metric.update per batch
        ops.pearson_stats(p[sl].contiguous(), t[sl].contiguous(), stats, layout="no")
    r, mean = ops.pearson_finalize(stats[0], want_mean=True)
    np.testing.assert_allclose(r.cpu().numpy(), g["scipy_r"], atol=1e-5)
    np.testing.assert_allclose(mean.cpu().numpy()[0], g["metric_pearson_mean"], atol=1e-5)
    # grouped (GroupedMetric, metrics/base.py:52-78)
    groups = torch.from_numpy(g["metric_groups"]).to(DEV)
    gstats = torch.zeros(3, 6, 17, device=DEV, dtype=torch.float64)
    ops.pearson_stats(p, t, gstats, layout="no", group=groups, n_groups=3)
    for key, val in zip(g["metric_grouped_keys"], g["metric_grouped_vals"]):
        _, m = ops.pearson_finalize(gstats[int(key)], want_mean=True)
        assert abs(float(m[0]) - val) < 1e-5


def test_pearson_bdt_layout_matches_scipy_loop():
    """main.py:459-477 on (b, d, t) tensors without the host rearrange: |dr| <= 1e-3 required, ~1e-6 expected."""
    torch.manual_seed(13)
    Bsz, D, T = 12, 1000, 100
    true = torch.randn(Bsz, D, T)
    pred = 0.2 * true + torch.randn(Bsz, D, T) + 0.5  # r ~ 0.2, non-zero mean
    ref = O.multidim_pearson_scipy(pred.numpy(), true.numpy())
    stats = torch.zeros(1, 6, D, device=DEV, dtype=torch.float64)
    ops.pearson_stats(pred.to(DEV), true.to(DEV), stats, layout="bdt")
    r, _ = ops.pearson_finalize(stats[0])
    assert float(np.abs(r.cpu().numpy() - ref).max()) < 1e-5
    # same numbers through the row-major fast path on the materialised (b t) x d matrices
    pf, tf = O.flatten_bdt(pred).contiguous().to(DEV), O.flatten_bdt(true).contiguous().to(DEV)
    stats2 = torch.zeros(1, 6, D, device=DEV, dtype=torch.float64)
    ops.pearson_stats(pf, tf, stats2, layout="no")
    r2, _ = ops.pearson_finalize(stats2[0])
    assert float(np.abs(r2.cpu().numpy() - ref).max()) < 1e-5
    np.testing.assert_allclose(r2.cpu().numpy(), O.pearson_columns_f64(pf.cpu().numpy(), tf.cpu().numpy()), atol=1e-5)
    # per-subject grouping on the bdt layout
    subj = torch.randint(0, 4, (Bsz,))
    gstats = torch.zeros(4, 6, D, device=DEV, dtype=torch.float64)
    ops.pearson_stats(pred.to(DEV), true.to(DEV), gstats, layout="bdt", group=subj.to(DEV), n_groups=4)
    for s in subj.unique().tolist():
        sel = subj == s
        ref_s = O.pearson_columns_f64(O.flatten_bdt(pred[sel]).numpy(), O.flatten_bdt(true[sel]).numpy())
        rs, _ = ops.pearson_finalize(gstats[s])
        assert float(np.abs(rs.cpu().numpy() - ref_s).max()) < 1e-5


@pytest.mark.parametrize("Bsz,D,T", [(1, 7, 100), (9, 1000, 100), (5, 37, 298), (6, 50, 33), (3, 11, 3), (2, 5, 1500)])
def test_pearson_bdt_shapes_slices_and_groups(Bsz, D, T):
    """(B, D, T) statistics kernel: vector (T % 4 == 0) and scalar item mapping, long-T fallback, parcel slices read in
    place (parcel-sharded evaluation), per-window group ids with repeated / alternating runs."""
    torch.manual_seed(15)
    true = torch.randn(Bsz, D, T)
    pred = 0.3 * true + torch.randn(Bsz, D, T) - 0.25
    ref = O.pearson_columns_f64(O.flatten_bdt(pred).numpy(), O.flatten_bdt(true).numpy())
    pd, td = pred.to(DEV), true.to(DEV)
    stats = torch.zeros(1, 6, D, device=DEV, dtype=torch.float64)
    ops.pearson_stats(pd, td, stats, layout="bdt")
    assert float(stats[0, 0].min()) == float(stats[0, 0].max()) == Bsz * T
    if Bsz * T > 2:
        r, _ = ops.pearson_finalize(stats[0])
        assert float(np.abs(r.cpu().numpy() - ref).max()) < 1e-5
    lo, hi = D // 3, D - D // 4
    st2 = torch.zeros(1, 6, hi - lo, device=DEV, dtype=torch.float64)
    ops.pearson_stats(pd[:, lo:hi], td[:, lo:hi], st2, layout="bdt")  # non-contiguous views, no copy
    torch.testing.assert_close(st2[0], stats[0][:, lo:hi], rtol=1e-6, atol=1e-4)  # block/atomic order differs
    groups = torch.tensor([(i // 2) % 3 for i in range(Bsz)])
    gst = torch.zeros(3, 6, D, device=DEV, dtype=torch.float64)
    ops.pearson_stats(pd, td, gst, layout="bdt", group=groups.to(DEV), n_groups=3)
    # group runs change how many windows share an fp32 partial sum before the fp64 fold: equal to fp32 rounding only
    torch.testing.assert_close(gst.sum(0), stats[0], rtol=1e-6, atol=1e-4)
    for s in groups.unique().tolist():
        sel = groups == s
        want = (pred[sel].double() * true[sel].double()).sum((0, 2))
        torch.testing.assert_close(gst[s, 5].cpu(), want, rtol=1e-6, atol=1e-4)


def test_pearson_ragged_and_odd_parcel_counts():
    torch.manual_seed(14)
    for n, o in ((1, 5), (7, 1003), (1000, 37), (4099, 1000)):
        p, t = torch.randn(n, o), torch.randn(n, o)
        stats = torch.zeros(1, 6, o, device=DEV, dtype=torch.float64)
        ops.pearson_stats(p.to(DEV), t.to(DEV), stats, layout="no")
        st = stats[0].cpu().numpy()
        np.testing.assert_allclose(st[0], n)
        np.testing.assert_allclose(st[1], p.double().sum(0).numpy(), rtol=1e-6, atol=1e-5)
        np.testing.assert_allclose(st[5], (p.double() * t.double()).sum(0).numpy(), rtol=1e-6, atol=1e-4)


# ------------------------------------------------------------------------------------------------- Pearson robustness
@pytest.mark.parametrize("layout", ["bdt", "no"])
def test_pearson_large_mean_constant_nan_and_tiny_n(layout):
    """scipy.stats.pearsonr (main.py:474-476) centres before it multiplies.  The pivoted statistics must agree with it
    to 1e-3 (north_star) when |mean| = 1e3 sigma, give NaN for constant columns and NaN inputs like scipy does, and
    +-1 for two rows; raw fp32 moments (shift=None, the round-1 kernel) are shown to FAIL the large-mean case, which is
    what the pivot is for."""
    import scipy.stats

    torch.manual_seed(21)
    Bsz, D, T = 6, 64, 100
    true = torch.randn(Bsz, D, T)
    pred = 0.3 * true + torch.randn(Bsz, D, T)
    true = true * 2.0 + 2000.0                      # mean = 1e3 sigma
    pred = pred * 0.5 - 500.0
    pred[:, 5] = 3.25                               # constant prediction column  -> scipy: NaN
    true[:, 9] = -7.0                               # constant target column
    pred[2, 11, 17] = float("nan")                  # NaN input                   -> NaN
    pf, tf = O.flatten_bdt(pred).contiguous(), O.flatten_bdt(true).contiguous()
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # float64 scipy on the same fp32-quantised inputs = the exact answer; float32 scipy (what main.py runs) for 1e-3
        ref = np.array([scipy.stats.pearsonr(tf[:, p].double().numpy(), pf[:, p].double().numpy())[0] for p in range(D)], dtype=np.float64)
        ref32 = np.array([scipy.stats.pearsonr(tf[:, p].numpy(), pf[:, p].numpy())[0] for p in range(D)], dtype=np.float64)
    if layout == "bdt":
        a, b = pred.to(DEV), true.to(DEV)
    else:
        a, b = pf.to(DEV), tf.to(DEV)
    r = ops.pearson_r(a, b, layout=layout)[0].cpu().numpy().astype(np.float64)
    nan_ref = np.isnan(ref)
    assert set(np.nonzero(nan_ref)[0]) == {5, 9, 11}
    assert np.array_equal(np.isnan(r), nan_ref), (np.nonzero(np.isnan(r))[0], np.nonzero(nan_ref)[0])
    assert float(np.abs(r[~nan_ref] - ref[~nan_ref]).max()) < 1e-5
    assert float(np.abs(r[~nan_ref] - ref32[~nan_ref]).max()) < 1e-3
    # the unpivoted moments lose the variance at this mean / sigma ratio
    raw = torch.zeros(1, 6, D, device=DEV, dtype=torch.float64)
    ops.pearson_stats(a, b, raw, layout=layout)
    r_raw = ops.pearson_finalize(raw[0])[0].cpu().numpy()
    ok = ~nan_ref & ~np.isnan(r_raw)
    assert (~np.isnan(r_raw[~nan_ref])).sum() == 0 or float(np.abs(r_raw[ok] - ref[ok]).max()) > 1e-3
    # streamed in three chunks with ONE pivot set + re-centred to pivot 0 (what the cross-rank merge does)
    shift = torch.empty(2, D, device=DEV, dtype=torch.float32)
    st = torch.zeros(1, 6, D, device=DEV, dtype=torch.float64)
    ops.pearson_pick_shift(a, b, shift, layout=layout)
    n0 = a.shape[0]
    for lo, hi in ((0, n0 // 3), (n0 // 3, n0 // 2), (n0 // 2, n0)):
        ops.pearson_stats(a[lo:hi].contiguous(), b[lo:hi].contiguous(), st, layout=layout, shift=shift)
    r_s = ops.pearson_finalize(st[0])[0].cpu().numpy()
    assert float(np.nanmax(np.abs(r_s - r))) < 1e-6
    ops.pearson_recenter(st, shift, None)
    want_sx = pf.double().sum(0).numpy()
    got = st[0].cpu().numpy()
    keep = ~nan_ref
    np.testing.assert_allclose(got[1][keep], want_sx[keep], rtol=1e-9)
    r_c = ops.pearson_finalize(st[0])[0].cpu().numpy()
    assert float(np.nanmax(np.abs(r_c[keep] - ref[keep]))) < 1e-4
    # two rows: r = +-1 exactly like scipy; one row: NaN (scipy raises)
    two_p, two_t = torch.tensor([[1000.5, -3.0], [1001.5, -4.0]]), torch.tensor([[7.0, 2.0], [9.0, 5.0]])
    r2 = ops.pearson_r(two_p.to(DEV), two_t.to(DEV), layout="no")[0].cpu().numpy()
    np.testing.assert_allclose(r2, [1.0, -1.0], atol=1e-6)
    r1 = ops.pearson_r(two_p[:1].contiguous().to(DEV), two_t[:1].contiguous().to(DEV), layout="no")[0].cpu().numpy()
    assert np.isnan(r1).all()


def test_pearson_metric_classes_use_pivots():
    from algonauts2025_b200.metrics import GroupedMetric, MultidimPearsonCorrCoef

    torch.manual_seed(22)
    true = torch.randn(8, 40, 100) + 3000.0
    pred = 0.4 * (true - 3000.0) + torch.randn(8, 40, 100) + 100.0
    ref = O.pearson_columns_f64(O.flatten_bdt(pred).numpy(), O.flatten_bdt(true).numpy())
    m = MultidimPearsonCorrCoef(num_outputs=40)
    for lo in range(0, 8, 3):
        m.update_bdt(pred[lo:lo + 3].to(DEV), true[lo:lo + 3].to(DEV))
    assert abs(float(m.compute()) - float(ref.mean())) < 1e-5
    assert float(np.abs(m.per_output().cpu().numpy() - ref).max()) < 1e-5
    m.reset()
    m.update(O.flatten_bdt(pred).contiguous().to(DEV), O.flatten_bdt(true).contiguous().to(DEV))
    assert abs(float(m.compute()) - float(ref.mean())) < 1e-5
    subj = torch.arange(8) % 2
    gm = GroupedMetric("MultidimPearsonCorrCoef", {"num_outputs": 40})
    gm.update_bdt(pred.to(DEV), true.to(DEV), groups=subj.to(DEV))
    got = gm.compute()
    for s in (0, 1):
        want = O.pearson_columns_f64(O.flatten_bdt(pred[subj == s]).numpy(), O.flatten_bdt(true[subj == s]).numpy()).mean()
        assert abs(got[str(s)] - want) < 1e-5
