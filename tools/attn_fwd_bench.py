"""Flash-style attention forward (tribe_attn_fwd: one launch) vs the two launches it replaces (tribe_attn_scores + the
batched P.V GEMM) at the train-step shape (16 x 8 heads x 298 tokens x 384 dims) and the eval batch (64 windows), cold L2.
Run under gpurun."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import algonauts2025_b200  # noqa: E402,F401
from algonauts2025_b200 import ops  # noqa: E402

dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)


def timeit(fn, iters=12):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.view(torch.int32).sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


def main():
    H, T, heads = 3072, 298, 8
    dh, Tp = H // heads, 304
    scale = dh ** -0.5
    for B in (16, 64):
        qkv = (torch.randn(B * T, 3 * H, device=dev) * 0.7).to(torch.bfloat16)
        P = torch.empty(B * heads, T, Tp, device=dev, dtype=torch.bfloat16)
        out = torch.empty(B * T, H, device=dev, dtype=torch.bfloat16)
        p_op = ops.Operand(P, inner=Tp, rows=T, row_stride=Tp, batch=B * heads, batch_stride=T * Tp)
        v_op = ops.Operand(qkv, inner=3 * H, rows=T, row_stride=3 * H, batch=B, batch_stride=T * 3 * H, mn_major=True, inner_off=2 * H, zin_stride=dh, zdiv=heads)

        def two():
            ops.attn_scores(qkv, 0, qkv, H, B, T, heads, dh, scale, P)
            ops.gemm(p_op, v_op, out, T, dh, Tp, ldd=H, batch=B * heads, z_inner=heads, d_zo=T * H, d_zi=dh)

        t2 = timeit(two)
        t1p = timeit(lambda: ops.attn_fwd(qkv, 0, H, 2 * H, B, T, heads, dh, scale, out, p_out=P))
        t1 = timeit(lambda: ops.attn_fwd(qkv, 0, H, 2 * H, B, T, heads, dh, scale, out))
        flops = 2 * 2 * B * heads * T * T * dh
        print(f"B = {B:3d}: scores + P.V (2 launches) {t2:7.1f} us | fused, P stored (training) {t1p:7.1f} us | fused, P not stored (inference) {t1:7.1f} us"
              f"   ({flops / t1p / 1e6:6.1f} / {flops / t1 / 1e6:6.1f} TFLOP/s useful)", flush=True)


if __name__ == "__main__":
    main()
