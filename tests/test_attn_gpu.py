"""Fused attention-score kernels (tribe_attn_scores): softmax(q k^T d^-1/2) and its backward formed in the tcgen05 epilogue,
against plain torch fp32 on the same bf16 operands.  Tolerances: bf16 outputs -> 1e-2 relative of the row scale."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import algonauts2025_b200  # noqa: E402,F401
from algonauts2025_b200 import ops  # noqa: E402

DEV = "cuda"


@pytest.mark.parametrize("Bsz,T,heads,dh", [(2, 74, 6, 64), (2, 298, 8, 384), (1, 160, 2, 128), (3, 161, 1, 64), (1, 320, 2, 64), (2, 7, 3, 64)])
def test_fused_scores_softmax_forward_and_backward(Bsz, T, heads, dh):
    torch.manual_seed(T + heads)
    H = heads * dh
    Tp = (T + 7) // 8 * 8
    qkv = (torch.randn(Bsz * T, 3 * H, device=DEV) * 1.5).to(torch.bfloat16)
    scale = dh ** -0.5
    P = torch.full((Bsz * heads, T, Tp), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.attn_scores(qkv, 0, qkv, H, Bsz, T, heads, dh, scale, P)
    q4 = qkv.float().view(Bsz, T, 3, heads, dh)
    S = torch.einsum("bihd,bjhd->bhij", q4[:, :, 0], q4[:, :, 1]) * scale
    ref_P = S.softmax(-1)
    got = P.float().view(Bsz, heads, T, Tp)
    assert torch.isfinite(got).all()
    assert float(got[..., T:].abs().max()) == 0.0 if Tp > T else True       # padding columns zeroed
    assert float((got[..., :T] - ref_P).abs().max()) <= 1e-2 * float(ref_P.max()) + 4e-3
    torch.testing.assert_close(got[..., :T].sum(-1), torch.ones(Bsz, heads, T, device=DEV), rtol=0, atol=2e-2)
    # backward: dS = P o (dP - rowsum(dP o P)) * scale with dP = dO V^T, P as stored (bf16)
    dO = (torch.randn(Bsz * T, H, device=DEV)).to(torch.bfloat16)
    dS = torch.full((Bsz * heads, T, Tp), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.attn_scores(dO, 0, qkv, 2 * H, Bsz, T, heads, dh, scale, dS, p_in=P)
    dP = torch.einsum("bihd,bjhd->bhij", dO.float().view(Bsz, T, heads, dh), q4[:, :, 2])
    Pf = got[..., :T]
    ref_dS = Pf * (dP - (dP * Pf).sum(-1, keepdim=True)) * scale
    gds = dS.float().view(Bsz, heads, T, Tp)
    assert torch.isfinite(gds).all()
    if Tp > T:
        assert float(gds[..., T:].abs().max()) == 0.0
    assert float((gds[..., :T] - ref_dS).abs().max()) <= 1e-2 * float(ref_dS.abs().max()) + 1e-4


def test_fused_attention_rejects_unsupported_shapes():
    x = torch.zeros(400, 3 * 64, device=DEV, dtype=torch.bfloat16)
    out = torch.zeros(1, 400, 400, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(algonauts2025_b200.TribeError):
        ops.attn_scores(x, 0, x, 64, 1, 400, 1, 64, 0.125, out)   # 400 keys do not fit one TMEM accumulator
    assert not ops.attn_fusable(400, 64) and not ops.attn_fusable(298, 96) and ops.attn_fusable(298, 384) == ops.FUSED_ATTN


@pytest.mark.parametrize("Bsz,T,heads,dh", [(2, 298, 8, 384), (1, 200, 2, 128), (1, 320, 2, 64), (2, 161, 1, 256), (3, 298, 2, 192)])
@pytest.mark.parametrize("keep_p", [True, False])
def test_flash_style_forward_matches_scores_then_pv(Bsz, T, heads, dh, keep_p):
    """tribe_attn_fwd: softmax(q k^T d^-1/2) v in ONE launch (S in TMEM, P as the shared-memory A operand of P.V) against
    (a) plain torch fp32 attention on the same bf16 operands and (b) the two-launch path it replaces (tribe_attn_scores +
    batched P.V GEMM): the P it optionally stores equals the scores kernel's up to one bf16 rounding step, O agrees to bf16 rounding."""
    torch.manual_seed(T + dh)
    H = heads * dh
    Tp = (T + 7) // 8 * 8
    assert ops.attn_fwd_supported(T, dh)
    qkv = (torch.randn(Bsz * T, 3 * H, device=DEV) * 1.2).to(torch.bfloat16)
    scale = dh ** -0.5
    out = torch.full((Bsz * T, H), float("nan"), device=DEV, dtype=torch.bfloat16)
    P = torch.full((Bsz * heads, T, Tp), float("nan"), device=DEV, dtype=torch.bfloat16) if keep_p else None
    ops.attn_fwd(qkv, 0, H, 2 * H, Bsz, T, heads, dh, scale, out, p_out=P)
    q4 = qkv.float().view(Bsz, T, 3, heads, dh)
    S = torch.einsum("bihd,bjhd->bhij", q4[:, :, 0], q4[:, :, 1]) * scale
    ref = torch.einsum("bhij,bjhd->bihd", S.softmax(-1), q4[:, :, 2]).reshape(Bsz * T, H)
    got = out.float()
    assert torch.isfinite(got).all()
    assert float((got - ref).abs().max()) <= 1.5e-2 * float(ref.abs().max())
    # the path it replaces
    P2 = torch.empty(Bsz * heads, T, Tp, device=DEV, dtype=torch.bfloat16)
    ops.attn_scores(qkv, 0, qkv, H, Bsz, T, heads, dh, scale, P2)
    out2 = torch.empty(Bsz * T, H, device=DEV, dtype=torch.bfloat16)
    p_op = ops.Operand(P2, inner=Tp, rows=T, row_stride=Tp, batch=Bsz * heads, batch_stride=T * Tp)
    v_op = ops.Operand(qkv, inner=3 * H, rows=T, row_stride=3 * H, batch=Bsz, batch_stride=T * 3 * H, mn_major=True, inner_off=2 * H,
                       zin_stride=dh, zdiv=heads)
    ops.gemm(p_op, v_op, out2, T, dh, Tp, ldd=H, batch=Bsz * heads, z_inner=heads, d_zo=T * H, d_zi=dh)
    if keep_p:
        # same exponentials; the row sum is formed as one chain here and as two half-row partial sums by the eight-warp
        # scores kernel, so single elements may differ by one bf16 rounding step
        assert float((P.float() - P2.float()).abs().max()) <= 2 ** -8 * float(P2.float().max())
        assert float((P != P2).float().mean()) < 0.05
    assert float((got - out2.float()).abs().max()) <= 8e-3 * float(ref.abs().max())


def test_flash_style_forward_rejects_unsupported_shapes():
    assert not ops.attn_fwd_supported(74, 64) and not ops.attn_fwd_supported(400, 64) and not ops.attn_fwd_supported(298, 96)
    assert ops.attn_fwd_supported(298, 384)
    x = torch.zeros(74, 3 * 64, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(algonauts2025_b200.TribeError):
        ops.attn_fwd(x, 0, 64, 128, 1, 74, 1, 64, 0.125, torch.zeros(74, 64, device=DEV, dtype=torch.bfloat16))
