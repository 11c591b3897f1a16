"""Bring-up check of the 2-CTA GEMM path (run under `timeout`): all operand majors, ragged shapes, epilogues, split-K."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import algonauts2025_b200  # noqa: E402
from algonauts2025_b200 import ops  # noqa: E402


def mn_op(x):
    mn, k = x.shape
    pad = (mn + 7) // 8 * 8
    store = torch.zeros(k, pad, device="cuda", dtype=x.dtype)
    store[:, :mn] = x.t()
    return ops.Operand(store, inner=mn, rows=k, row_stride=pad, mn_major=True)


ok = True
for (m, n, k) in ((1024, 512, 256), (1100, 2900, 1000), (4768, 3072, 3072), (4768, 3072, 12288)):
    for a_mn in (False, True):
        for b_mn in (False, True):
            torch.manual_seed(0)
            A = (torch.randn(m, k, device="cuda") / math.sqrt(k)).bfloat16()
            B = torch.randn(n, k, device="cuda").bfloat16()
            ref = A.float() @ B.float().t()
            out = torch.full((m, n), float("nan"), device="cuda")
            ops.gemm(mn_op(A) if a_mn else ops.kmajor(A), mn_op(B) if b_mn else ops.kmajor(B), out, m, n, k, ldd=n, block_n=256)
            torch.cuda.synchronize()
            err = (out - ref).abs()
            bad = float((~(err <= 1e-2 * float(ref.abs().max()) + 1e-2 * ref.abs())).float().mean())
            print(f"m{m} n{n} k{k} A{'mn' if a_mn else 'k'} B{'mn' if b_mn else 'k'}: max_err {float(err.nan_to_num(1e9).max()):.3e} frac_bad {bad:.4f}", flush=True)
            ok &= bad == 0.0
print("2CTA", "OK" if ok else "FAILED")
