"""Small driver for ncu captures of the bandwidth kernels (pool, sub-layer tail, Pearson, Adam, ingest)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import algonauts2025_b200  # noqa: E402
from algonauts2025_b200 import _lib, ops  # noqa: E402
import ctypes  # noqa: E402

dev = "cuda"
x = torch.randn(256, 1000, 298, device=dev)
dy = torch.randn(256, 1000, 100, device=dev)
M, H = 4768, 3072
xs, dyo, dxn = torch.randn(M, H, device=dev), torch.randn(M, H, device=dev), torch.randn(M, H, device=dev).bfloat16()
rn, g, rs = torch.rand(M, device=dev), torch.ones(1, device=dev), torch.ones(H, device=dev)
dx, dxb, d_rs, d_g = torch.empty(M, H, device=dev), torch.empty(M, H, device=dev, dtype=torch.bfloat16), torch.zeros(H, device=dev), torch.zeros(1, device=dev)
pred, true = torch.randn(64000, 1000, device=dev), torch.randn(64000, 1000, device=dev)
stats = torch.zeros(1, 6, 1000, device=dev, dtype=torch.float64)
feats = torch.randn(16, 2, 3072, 298, device=dev)
fout = torch.empty(M, 6144, device=dev, dtype=torch.bfloat16)
n = 100_000_000
p, gr, m, v = (torch.randn(n, device=dev) for _ in range(4))
v.abs_()
p16 = torch.empty(n, device=dev, dtype=torch.bfloat16)
lib = _lib.load()
for _ in range(2):
    ops.adaptive_avg_pool_fwd(x, 100)
    ops.adaptive_avg_pool_bwd(dy, 298)
    ops.sublayer_bwd(dyo, dxn, xs, rn, g, rs, dx, dxb, d_rs, d_g)
    ops.pearson_stats(pred, true, stats, layout="no")
    ops.ingest_features(feats, fout, 0, False)
    lib.tribe_adam_step(ctypes.c_void_p(p.data_ptr()), ctypes.c_void_p(gr.data_ptr()), ctypes.c_void_p(m.data_ptr()), ctypes.c_void_p(v.data_ptr()),
                        ctypes.c_void_p(p16.data_ptr()), n, 1e-4, 0.9, 0.999, 1e-8, 0.0, 3, 0, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("ok")
