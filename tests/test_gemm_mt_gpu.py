"""B-stationary multi-row-tile GEMM (csrc/gemm_mt_sm100.cuh) on the attention contractions of the encoder
(x_transformers attention behind /root/reference/algonauts2025/model.py:173, restated in oracle/xt_encoder.py:85-94):
O = P V, dV = P^T dO, dQ = dS K (+ inverse rotary), dK = dS^T Q — vs fp32 torch and vs the generic 1-CTA kernel
(``block_n`` given = generic path), all four operand-major combinations, 2 and 3 row tiles, ragged N."""
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    import algonauts2025_b200  # noqa: F401
    from algonauts2025_b200 import ops
    return ops


def bf(x):
    return x.to(torch.bfloat16)


def close(a, b, tol=2e-2):
    a, b = a.float(), b.float()
    err = (a - b).abs().max().item()
    ref = b.abs().max().item() + 1e-6
    assert err <= tol * ref, f"max err {err} vs scale {ref}"


@pytest.mark.parametrize("Bsz,T,H,dh", [(2, 298, 8, 384), (3, 200, 2, 128), (1, 257, 2, 256)])
def test_attention_contractions_mt_vs_torch_and_generic(Bsz, T, H, dh):
    ops = _ops()
    torch.manual_seed(11)
    dim, Tp = H * dh, (T + 7) // 8 * 8
    BH = Bsz * H
    qkv = bf(torch.randn(Bsz * T, 3 * dim, device=DEV))
    dO = bf(torch.randn(Bsz * T, dim, device=DEV))
    P = torch.zeros(BH, T, Tp, device=DEV, dtype=torch.bfloat16)
    P[..., :T] = bf(torch.rand(BH, T, T, device=DEV).softmax(-1) * 4)
    dS = torch.zeros_like(P)
    dS[..., :T] = bf(torch.randn(BH, T, T, device=DEV) * 0.1)
    q4 = qkv.float().view(Bsz, T, 3, H, dh)
    Pf = P.float().view(Bsz, H, T, Tp)[..., :T]
    dSf = dS.float().view(Bsz, H, T, Tp)[..., :T]
    dOf = dO.float().view(Bsz, T, H, dh)

    def operands():
        p_op = ops.Operand(P, inner=Tp, rows=T, row_stride=Tp, batch=BH, batch_stride=T * Tp)
        v_op = ops.Operand(qkv, inner=3 * dim, rows=T, row_stride=3 * dim, batch=Bsz, batch_stride=T * 3 * dim, mn_major=True,
                           inner_off=2 * dim, zin_stride=dh, zdiv=H)
        pt_op = ops.Operand(P, inner=Tp, rows=T, row_stride=Tp, batch=BH, batch_stride=T * Tp, mn_major=True)
        dom_op = ops.Operand(dO, inner=dim, rows=T, row_stride=dim, batch=Bsz, batch_stride=T * dim, mn_major=True, zin_stride=dh, zdiv=H)
        ds_op = ops.Operand(dS, inner=Tp, rows=T, row_stride=Tp, batch=BH, batch_stride=T * Tp)
        km_op = ops.Operand(qkv, inner=3 * dim, rows=T, row_stride=3 * dim, batch=Bsz, batch_stride=T * 3 * dim, mn_major=True,
                            inner_off=dim, zin_stride=dh, zdiv=H)
        dst_op = ops.Operand(dS, inner=Tp, rows=T, row_stride=Tp, batch=BH, batch_stride=T * Tp, mn_major=True)
        qm_op = ops.Operand(qkv, inner=3 * dim, rows=T, row_stride=3 * dim, batch=Bsz, batch_stride=T * 3 * dim, mn_major=True,
                            zin_stride=dh, zdiv=H)
        return p_op, v_op, pt_op, dom_op, ds_op, km_op, dst_op, qm_op

    def run(block_n):
        p_op, v_op, pt_op, dom_op, ds_op, km_op, dst_op, qm_op = operands()
        attn = torch.full((Bsz * T, dim), float("nan"), device=DEV, dtype=torch.bfloat16)
        dqkv = torch.full((Bsz * T, 3 * dim), float("nan"), device=DEV, dtype=torch.bfloat16)
        kw = dict(batch=BH, z_inner=H, d_zi=dh, block_n=block_n)
        ops.gemm(p_op, v_op, attn, T, dh, Tp, ldd=dim, d_zo=T * dim, **kw)                       # O  = P V
        ops.gemm(pt_op, dom_op, dqkv, T, dh, T, ldd=3 * dim, d_zo=T * 3 * dim, d_off=2 * dim, **kw)  # dV = P^T dO
        ops.gemm(ds_op, km_op, dqkv, T, dh, Tp, ldd=3 * dim, d_zo=T * 3 * dim, **kw)                 # dQ = dS K
        ops.gemm(dst_op, qm_op, dqkv, T, dh, T, ldd=3 * dim, d_zo=T * 3 * dim, d_off=dim, **kw)      # dK = dS^T Q
        torch.cuda.synchronize()
        return attn, dqkv

    attn, dqkv = run(0)          # multi-row-tile path (n = dh is a multiple of 128, 2-3 row tiles, <= 8 k-blocks)
    attn_g, dqkv_g = run(128)    # generic 1-CTA kernel
    ref_O = torch.einsum("bhij,bjhd->bihd", Pf, q4[:, :, 2]).reshape(Bsz * T, dim)
    ref_dV = torch.einsum("bhij,bihd->bjhd", Pf, dOf).reshape(Bsz * T, dim)
    ref_dQ = torch.einsum("bhij,bjhd->bihd", dSf, q4[:, :, 1]).reshape(Bsz * T, dim)
    ref_dK = torch.einsum("bhij,bihd->bjhd", dSf, q4[:, :, 0]).reshape(Bsz * T, dim)
    close(attn, ref_O)
    close(dqkv[:, 2 * dim:], ref_dV)
    close(dqkv[:, :dim], ref_dQ)
    close(dqkv[:, dim:2 * dim], ref_dK)
    # same products, same fp32 accumulation order along K inside a tile -> identical bf16 results
    assert torch.equal(attn, attn_g)
    assert torch.equal(dqkv, dqkv_g)


def test_mt_kmajor_b_alpha_f32_out_and_rope_epilogue():
    """The remaining operand-major combinations (K-major B; MN-major A with K-major B), fp32 output with alpha, and the
    inverse-rotary epilogue of dQ / dK."""
    ops = _ops()
    torch.manual_seed(12)
    Z, M, N, K = 6, 296, 256, 200
    A = bf(torch.randn(Z, M, K, device=DEV))
    Bm = bf(torch.randn(Z, N, K, device=DEV))
    out = torch.full((Z, M, N), float("nan"), device=DEV)
    a_op = ops.Operand(A, inner=K, rows=M, row_stride=K, batch=Z, batch_stride=M * K)
    b_op = ops.Operand(Bm, inner=K, rows=N, row_stride=K, batch=Z, batch_stride=N * K)
    ops.gemm(a_op, b_op, out, M, N, K, ldd=N, batch=Z, d_zo=M * N, alpha=0.5)
    close(out, 0.5 * torch.einsum("zmk,znk->zmn", A.float(), Bm.float()), tol=1e-2)
    At = A.transpose(1, 2).contiguous()  # (Z, K, M): MN-major A
    at_op = ops.Operand(At, inner=M, rows=K, row_stride=M, batch=Z, batch_stride=K * M, mn_major=True)
    out2 = torch.full((Z, M, N), float("nan"), device=DEV)
    ops.gemm(at_op, b_op, out2, M, N, K, ldd=N, batch=Z, d_zo=M * N, alpha=0.5)
    assert torch.equal(out, out2)
    # rotary epilogue: compare against the generic kernel bit for bit
    T, dh, rot = M, 128, 64
    pos = torch.arange(T, device=DEV, dtype=torch.float32)[:, None]
    inv = 1.0 / (10000 ** (torch.arange(0, rot, 2, device=DEV, dtype=torch.float32) / rot))
    ang = pos * inv[None]
    table = torch.stack([ang.cos(), ang.sin()], -1).contiguous()  # (T, rot/2, 2)
    res = []
    for bn in (0, 128):
        o = torch.full((Z, M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
        ops.gemm(a_op, b_op, o, M, N, K, ldd=N, batch=Z, d_zo=M * N, epilogue=ops.EPI_ROPE, rope=table, rope_t=T, rope_dim=rot, head_dim=dh,
                 rope_cols=N, rope_sign=-1.0, block_n=bn)
        res.append(o)
    torch.cuda.synchronize()
    assert torch.equal(res[0], res[1])
    assert not torch.isnan(res[0].float()).any()
