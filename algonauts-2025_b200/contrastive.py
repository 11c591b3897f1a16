"""Contrastive alignment branch (reference algonauts2025/model.py:177-241, enabled by default in
algonauts2025/grids/defaults.py:102): modality latents through the contrastive head and the symmetric InfoNCE, as
autograd Functions over the sm_100a kernels.

InfoNCE on B200: rows are L2-normalised by the ScaleNorm kernel (g = 1/sqrt(H) makes it a plain normalise), the
(n x n) logits are ONE tcgen05 GEMM, the two cross-entropies are a single pass over the logits (exp-sums by row and by
column with the constant shift 1/tau) and the gradient is one elementwise pass + two GEMMs (G k^ and G^T q^)."""
from __future__ import annotations

import torch

from . import ops
from ._lib import TribeError


def _round8(n: int) -> int:
    return (n + 7) // 8 * 8


class _HeadFn(torch.autograd.Function):
    """``contrastive_heads[modality](layer-aggregated features)`` -> (B, T, H) fp32 (model.py:199-206)."""

    @staticmethod
    def forward(ctx, anchor, model, data, modality):
        eng = model._engine
        eng._check_flat()
        eng.flat.refresh_bf16()
        lin = model.contrastive_heads[modality]
        x = data.to(eng.device, non_blocking=True)
        if x.dim() == 3:
            x = x.unsqueeze(1)
        B, T = x.shape[0], x.shape[-1]
        M, K, H = B * T, lin.in_features, lin.out_features
        k_in = x.shape[2] if model.config.layer_aggregation == "mean" else x.shape[1] * x.shape[2]
        if k_in != K:
            raise TribeError(f"contrastive head '{modality}': feature width {k_in} does not match in_features {K}")
        feat = torch.empty(M, K, device=eng.device, dtype=torch.bfloat16)
        ops.ingest_features(x, feat, 0, model.config.layer_aggregation == "mean")
        out = torch.empty(M, H, device=eng.device, dtype=torch.float32)
        w16 = eng._w16(f"contrastive_heads.{modality}.weight")
        ops.gemm(ops.kmajor(feat), ops.kmajor(w16), out, M, H, K, ldd=H, bias=eng._p(f"contrastive_heads.{modality}.bias"))
        ctx.model, ctx.modality, ctx.feat, ctx.dims = model, modality, feat, (M, K, H)
        return out.view(B, T, H)

    @staticmethod
    def backward(ctx, grad_out):
        model, modality, feat = ctx.model, ctx.modality, ctx.feat
        M, K, H = ctx.dims
        eng = model._engine
        fl = eng.flat
        fl.ensure_grad()
        acc = {}
        wn, bn = f"contrastive_heads.{modality}.weight", f"contrastive_heads.{modality}.bias"
        gw, gb = eng._grad_target(wn, acc), eng._grad_target(bn, acc)
        go = grad_out.contiguous().float().view(M, H)
        gob = torch.empty(M, H, device=eng.device, dtype=torch.bfloat16)
        ops.cast_f32_bf16(go.view(-1), gob.view(-1))
        eng._wgrad(ops.mnmajor(gob), ops.mnmajor(feat), wn, H, K, M, acc[wn])
        ops.colsum(go, gb, accumulate=acc[bn])
        for n in (wn, bn):
            p = fl.params[n]
            if p.grad is None:
                p.grad = fl.gview(n)
        if eng.comm is not None:
            eng.comm.bucket_ready(0)
        return None, None, None, None


def modality_latents(model, batch, modality: str) -> torch.Tensor:
    data = batch.data[modality]
    needs_grad = torch.is_grad_enabled() and model.contrastive_heads[modality].weight.requires_grad
    if needs_grad:
        return _HeadFn.apply(model._anchor(), model, data, modality)  # fresh leaf, see FmriEncoder._anchor

    class _Ctx:  # no-grad path reuses the forward body
        pass

    return _HeadFn.forward(_Ctx(), None, model, data, modality)


class _NceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, tau):
        if not q.is_cuda or not k.is_cuda:
            raise TribeError("InfoNCE needs CUDA tensors (no CPU fallback)")
        h = q.shape[-1]
        q2, k2 = q.detach().reshape(-1, h).float().contiguous(), k.detach().reshape(-1, h).float().contiguous()
        n = q2.shape[0]
        dev = q.device
        ginv = torch.full((1,), h ** -0.5, device=dev, dtype=torch.float32)  # sqrt(H) * g = 1 -> F.normalize
        qh, kh = torch.empty(n, h, device=dev, dtype=torch.bfloat16), torch.empty(n, h, device=dev, dtype=torch.bfloat16)
        rq, rk = torch.empty(n, device=dev), torch.empty(n, device=dev)
        ops.scalenorm_fwd(q2, ginv, qh, rq)
        ops.scalenorm_fwd(k2, ginv, kh, rk)
        ld = _round8(n)
        logits = torch.empty(n, ld, device=dev, dtype=torch.float32)
        ops.gemm(ops.kmajor(qh), ops.kmajor(kh), logits, n, n, h, ldd=ld, alpha=1.0 / tau)
        shift = 1.0 / tau
        row_sum, col_sum = torch.empty(n, device=dev), torch.empty(n, device=dev)
        ops.nce_expsums(logits, n, shift, row_sum, col_sum)
        loss = torch.empty(1, device=dev, dtype=torch.float32)
        ops.nce_loss(logits, n, shift, row_sum, col_sum, loss)
        ctx.save_for_backward(q2, k2, qh, kh, rq, rk, logits, row_sum, col_sum, ginv)
        ctx.meta = (n, h, ld, tau, q.shape, k.shape)
        return loss[0]

    @staticmethod
    def backward(ctx, grad):
        q2, k2, qh, kh, rq, rk, logits, row_sum, col_sum, ginv = ctx.saved_tensors
        n, h, ld, tau, qshape, kshape = ctx.meta
        dev = q2.device
        g = torch.empty(n, ld, device=dev, dtype=torch.bfloat16)
        up = grad.detach().float().reshape(1).contiguous()
        ops.nce_grad(logits, n, 1.0 / tau, row_sum, col_sum, up, 0.5 / (n * tau), g)
        dqh, dkh = torch.empty(n, h, device=dev, dtype=torch.bfloat16), torch.empty(n, h, device=dev, dtype=torch.bfloat16)
        # dq^ = G k^ (K-major A over the key index, k^ read MN-major);  dk^ = G^T q^ (both MN-major)
        ops.gemm(ops.Operand(g, inner=ld, rows=n, row_stride=ld), ops.Operand(kh, inner=h, rows=n, row_stride=h, mn_major=True), dqh, n, h, ld, ldd=h)
        ops.gemm(ops.Operand(g, inner=ld, rows=n, row_stride=ld, mn_major=True), ops.Operand(qh, inner=h, rows=n, row_stride=h, mn_major=True), dkh,
                 n, h, n, ldd=h)
        dq, dk = torch.empty(n, h, device=dev), torch.empty(n, h, device=dev)
        ops.sublayer_bwd(None, dqh, q2, rq, ginv, None, dq, None, None, None)  # backward of F.normalize
        ops.sublayer_bwd(None, dkh, k2, rk, ginv, None, dk, None, None, None)
        return dq.view(qshape), dk.view(kshape), None


def info_nce(q: torch.Tensor, k: torch.Tensor, tau: float = 0.07) -> torch.Tensor:
    """Symmetric InfoNCE over flattened [B, T, H] sequences (model.py:208-221)."""
    return _NceFn.apply(q, k, tau)
