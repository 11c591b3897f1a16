"""One forward (mode 0) and one backward (mode 1) launch of the attention-scores kernel + the four batched contractions at
the train-step shape — the target of an `ncu --set full` capture (see profiles/)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import algonauts2025_b200  # noqa: E402,F401
from algonauts2025_b200 import ops  # noqa: E402

B, T, heads, dh = 16, 298, 8, 384
H, Tp, BH = heads * dh, 304, 16 * 8
dev = "cuda"
qkv = (torch.randn(B * T, 3 * H, device=dev) * 0.5).bfloat16()
dO = torch.randn(B * T, H, device=dev).bfloat16()
P = torch.zeros(BH, T, Tp, device=dev, dtype=torch.bfloat16)
dS = torch.zeros_like(P)
out = torch.empty(B * T, H, device=dev, dtype=torch.bfloat16)
scale = dh ** -0.5
for _ in range(2):
    ops.attn_scores(qkv, 0, qkv, H, B, T, heads, dh, scale, P)
    ops.attn_scores(dO, 0, qkv, 2 * H, B, T, heads, dh, scale, dS, p_in=P)
    p_op = ops.Operand(P, inner=Tp, rows=T, row_stride=Tp, batch=BH, batch_stride=T * Tp)
    v_op = ops.Operand(qkv, inner=3 * H, rows=T, row_stride=3 * H, batch=B, batch_stride=T * 3 * H, mn_major=True, inner_off=2 * H, zin_stride=dh, zdiv=heads)
    ops.gemm(p_op, v_op, out, T, dh, Tp, ldd=H, batch=BH, z_inner=heads, d_zo=T * H, d_zi=dh)
    torch.cuda.synchronize()
print("ok")
