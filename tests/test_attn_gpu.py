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
