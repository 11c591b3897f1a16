"""Fixed vs per-tile cost of the batched attention contractions: P.V (the 1-CTA kernel, 128 x 192 tiles, 5 k-blocks each)
timed at 1/4 ... 4x the train-step batch.  A straight line time = a + b * tiles separates what a launch costs
(launch + pipeline fill + last drain) from what a tile costs.  Run under gpurun."""
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


def timeit(fn, iters=12, warm=False):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if not warm:
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
    rows = []
    for B in (2, 4, 8, 16, 32, 64):
        qkv = (torch.randn(B * T, 3 * H, device=dev) * 0.05).to(torch.bfloat16)
        P = (torch.rand(B * HEADS, T, TP, device=dev)).to(torch.bfloat16)
        out = torch.empty(B * T, H, device=dev, dtype=torch.bfloat16)
        p_op = ops.Operand(P, inner=TP, rows=T, row_stride=TP, batch=B * HEADS, batch_stride=T * TP)
        v_op = ops.Operand(qkv, inner=3 * H, rows=T, row_stride=3 * H, batch=B, batch_stride=T * 3 * H, mn_major=True, inner_off=2 * H, zin_stride=DH, zdiv=HEADS)
        fn = lambda: ops.gemm(p_op, v_op, out, T, DH, TP, ldd=H, batch=B * HEADS, z_inner=HEADS, d_zo=T * H, d_zi=DH)  # noqa: E731
        tiles = B * HEADS * 3 * 2
        cold, warm = timeit(fn), timeit(fn, warm=True)
        rows.append((tiles, cold, warm))
        print(f"P.V  batch {B:3d} x {HEADS} heads: {tiles:5d} tiles ({tiles / 148:5.2f} per CTA)   cold L2 {cold:7.1f} us   warm L2 {warm:7.1f} us", flush=True)
    # least squares on the multi-wave points
    import numpy as np

    for name, col in (("cold", 1), ("warm", 2)):
        pts = [(r[0] / 148.0, r[col]) for r in rows if r[0] >= 148]
        A = np.array([[1.0, x] for x, _ in pts])
        y = np.array([t for _, t in pts])
        a, b = np.linalg.lstsq(A, y, rcond=None)[0]
        print(f"{name} L2: time ~= {a:.1f} us per launch + {b:.2f} us per tile-wave (one 128 x 192 x 320 tile per CTA; its MMAs alone take ~1 us)")
    # an empty-ish launch of the same kernel: one tile per CTA
    print("(the smallest case above has < 1 tile per CTA: its time is the launch floor of this kernel)")


if __name__ == "__main__":
    main()
