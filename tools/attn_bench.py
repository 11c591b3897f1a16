"""Fused attention-score kernels vs the generic GEMM + softmax path at the TRIBE shape (run under gpurun)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import algonauts2025_b200  # noqa
from algonauts2025_b200 import ops

B, T, heads, dh = 16, 298, 8, 384
H, Tp, BH = heads * dh, 304, 16 * 8
dev = "cuda"
qkv = torch.randn(B * T, 3 * H, device=dev).bfloat16()
dO = torch.randn(B * T, H, device=dev).bfloat16()
S = torch.empty(BH, T, Tp, device=dev)
P = torch.empty(BH, T, Tp, device=dev, dtype=torch.bfloat16)
dS = torch.empty_like(P)
scale = dh ** -0.5
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def timeit(name, fn, iters=20):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print(f"{name:50s} {1e3 * ts[len(ts) // 2]:8.1f} us", flush=True)


def old_fwd():
    q = ops.Operand(qkv, inner=3 * H, rows=T, row_stride=3 * H, batch=B, batch_stride=T * 3 * H, zin_stride=dh, zdiv=heads)
    k = ops.Operand(qkv, inner=3 * H, rows=T, row_stride=3 * H, batch=B, batch_stride=T * 3 * H, inner_off=H, zin_stride=dh, zdiv=heads)
    ops.gemm(q, k, S, T, T, dh, ldd=Tp, batch=BH, z_inner=heads, d_zo=heads * T * Tp, d_zi=T * Tp, alpha=scale)
    ops.softmax_fwd(S, P, T)


def old_bwd():
    do = ops.Operand(dO, inner=H, rows=T, row_stride=H, batch=B, batch_stride=T * H, zin_stride=dh, zdiv=heads)
    v = ops.Operand(qkv, inner=3 * H, rows=T, row_stride=3 * H, batch=B, batch_stride=T * 3 * H, inner_off=2 * H, zin_stride=dh, zdiv=heads)
    ops.gemm(do, v, S, T, T, dh, ldd=Tp, batch=BH, z_inner=heads, d_zo=heads * T * Tp, d_zi=T * Tp)
    ops.softmax_bwd(P, S, dS, scale, T)


timeit("generic  S=QK^T (fp32 to HBM) + softmax kernel", old_fwd)
timeit("fused    softmax(QK^T) in the tcgen05 epilogue", lambda: ops.attn_scores(qkv, 0, qkv, H, B, T, heads, dh, scale, P))
timeit("generic  dP=dO V^T (fp32 to HBM) + softmax-bwd kernel", old_bwd)
timeit("fused    dS in the tcgen05 epilogue", lambda: ops.attn_scores(dO, 0, qkv, 2 * H, B, T, heads, dh, scale, dS, p_in=P))

# ---- does the packed (M, 3H) layout cost anything?  Same contractions on per-head contiguous copies (BH, T, dh)
q4 = qkv.view(B, T, 3, heads, dh)
qc = q4[:, :, 0].permute(0, 2, 1, 3).contiguous().view(BH, T, dh)
kc = q4[:, :, 1].permute(0, 2, 1, 3).contiguous().view(BH, T, dh)


def contiguous_fwd():
    q = ops.Operand(qc, inner=dh, rows=T, row_stride=dh, batch=BH, batch_stride=T * dh)
    k = ops.Operand(kc, inner=dh, rows=T, row_stride=dh, batch=BH, batch_stride=T * dh)
    ops.gemm(q, k, S, T, T, dh, ldd=Tp, batch=BH, d_zo=T * Tp, alpha=scale)


def packed_fwd():
    q = ops.Operand(qkv, inner=3 * H, rows=T, row_stride=3 * H, batch=B, batch_stride=T * 3 * H, zin_stride=dh, zdiv=heads)
    k = ops.Operand(qkv, inner=3 * H, rows=T, row_stride=3 * H, batch=B, batch_stride=T * 3 * H, inner_off=H, zin_stride=dh, zdiv=heads)
    ops.gemm(q, k, S, T, T, dh, ldd=Tp, batch=BH, z_inner=heads, d_zo=heads * T * Tp, d_zi=T * Tp, alpha=scale)


timeit("S=QK^T only, operands inside packed qkv (M, 3H)", packed_fwd)
timeit("S=QK^T only, per-head contiguous operands", contiguous_fwd)
Sb = torch.empty(BH, T, Tp, device=dev, dtype=torch.bfloat16)


def packed_fwd_bf16():
    q = ops.Operand(qkv, inner=3 * H, rows=T, row_stride=3 * H, batch=B, batch_stride=T * 3 * H, zin_stride=dh, zdiv=heads)
    k = ops.Operand(qkv, inner=3 * H, rows=T, row_stride=3 * H, batch=B, batch_stride=T * 3 * H, inner_off=H, zin_stride=dh, zdiv=heads)
    ops.gemm(q, k, Sb, T, T, dh, ldd=Tp, batch=BH, z_inner=heads, d_zo=heads * T * Tp, d_zi=T * Tp, alpha=scale)


timeit("S=QK^T only, packed operands, bf16 output", packed_fwd_bf16)
# warm-L2 variants (no flush between launches): what the step sees right after the QKV GEMM wrote qkv
for name, fn in (("warm L2: packed fp32 out", packed_fwd), ("warm L2: contiguous fp32 out", contiguous_fwd),
                 ("warm L2: fused softmax", lambda: ops.attn_scores(qkv, 0, qkv, H, B, T, heads, dh, scale, P))):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        fn()
    b.record(); torch.cuda.synchronize()
    print(f"{name:50s} {1e3 * a.elapsed_time(b) / 20:8.1f} us", flush=True)
