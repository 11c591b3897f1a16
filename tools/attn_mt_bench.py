"""The four batched attention contractions of a layer (O = P V, dV = P^T dO, dQ = dS K, dK = dS^T Q) at the train-step shape:
B-stationary multi-row-tile kernel (gemm_mt_sm100.cuh, default) vs the generic 1-CTA kernel (block_n = 192: what ran
before).  Cold L2 (256 MB read between launches), median of 15.  Run under gpurun."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import algonauts2025_b200  # noqa: E402,F401
from algonauts2025_b200 import ops  # noqa: E402

H, T, HEADS = 3072, 298, 8
DH, TP = H // HEADS, 304
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)


def timeit(fn, iters=15):
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
    for B in (16, 64):
        BH = B * HEADS
        qkv = (torch.randn(B * T, 3 * H, device=dev) * 0.05).to(torch.bfloat16)
        dO = (torch.randn(B * T, H, device=dev) * 0.05).to(torch.bfloat16)
        P = torch.rand(BH, T, TP, device=dev).to(torch.bfloat16)
        out = torch.empty(B * T, H, device=dev, dtype=torch.bfloat16)
        dqkv = torch.empty(B * T, 3 * H, device=dev, dtype=torch.bfloat16)
        p_op = ops.Operand(P, inner=TP, rows=T, row_stride=TP, batch=BH, batch_stride=T * TP)
        pt_op = ops.Operand(P, inner=TP, rows=T, row_stride=TP, batch=BH, batch_stride=T * TP, mn_major=True)
        v_op = ops.Operand(qkv, inner=3 * H, rows=T, row_stride=3 * H, batch=B, batch_stride=T * 3 * H, mn_major=True, inner_off=2 * H, zin_stride=DH, zdiv=HEADS)
        k_op = ops.Operand(qkv, inner=3 * H, rows=T, row_stride=3 * H, batch=B, batch_stride=T * 3 * H, mn_major=True, inner_off=H, zin_stride=DH, zdiv=HEADS)
        q_op = ops.Operand(qkv, inner=3 * H, rows=T, row_stride=3 * H, batch=B, batch_stride=T * 3 * H, mn_major=True, zin_stride=DH, zdiv=HEADS)
        do_op = ops.Operand(dO, inner=H, rows=T, row_stride=H, batch=B, batch_stride=T * H, mn_major=True, zin_stride=DH, zdiv=HEADS)
        kw = dict(batch=BH, z_inner=HEADS, d_zi=DH)
        cases = {
            "O  = P V     (A K-major,  B MN-major)": lambda bn: ops.gemm(p_op, v_op, out, T, DH, TP, ldd=H, d_zo=T * H, block_n=bn, **kw),
            "dV = P^T dO  (A MN-major, B MN-major)": lambda bn: ops.gemm(pt_op, do_op, dqkv, T, DH, T, ldd=3 * H, d_zo=T * 3 * H, d_off=2 * H, block_n=bn, **kw),
            "dQ = dS K    (A K-major,  B MN-major)": lambda bn: ops.gemm(p_op, k_op, dqkv, T, DH, TP, ldd=3 * H, d_zo=T * 3 * H, block_n=bn, **kw),
            "dK = dS^T Q  (A MN-major, B MN-major)": lambda bn: ops.gemm(pt_op, q_op, dqkv, T, DH, T, ldd=3 * H, d_zo=T * 3 * H, d_off=H, block_n=bn, **kw),
        }
        flops = 2.0 * BH * T * T * DH
        for name, fn in cases.items():
            t_mt = timeit(lambda: fn(0))
            t_gen = timeit(lambda: fn(192))
            print(f"B={B:3d} {name}:  multi-row-tile {t_mt:7.1f} us ({flops / t_mt / 1e6:6.0f} TFLOP/s)   generic 128x192 {t_gen:7.1f} us ({flops / t_gen / 1e6:6.0f} TFLOP/s)",
                  flush=True)


if __name__ == "__main__":
    main()
