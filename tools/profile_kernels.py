"""Workloads for `ncu --set full` captures (run under gpurun, one GPU; see profiles/ for the summaries).

    python tools/profile_kernels.py bw     # Pearson (b,d,t) statistics at the eval-sweep size, sublayer_bwd, pearson-loss bwd
    python tools/profile_kernels.py step   # two eager full-size train steps (every GEMM variant, attention, Adam)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import algonauts2025_b200  # noqa: E402
from algonauts2025_b200 import ops  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "bw"
dev = "cuda"
if mode == "bw":
    O, TQ = 1000, 100
    p, t = torch.randn(2560, O, TQ, device=dev), torch.randn(2560, O, TQ, device=dev)
    stats = torch.zeros(1, 6, O, device=dev, dtype=torch.float64)
    for _ in range(2):
        ops.pearson_stats(p, t, stats, layout="bdt")
    _, coef = ops.pearson_loss_fwd(p, t, layout="bdt")
    up = torch.ones(1, device=dev)
    ops.pearson_loss_bwd(p, t, coef, up, layout="bdt")
    M, H = 4768, 3072
    xs = torch.randn(M, H, device=dev)
    g = torch.ones(1, device=dev)
    y, rn = torch.empty(M, H, device=dev, dtype=torch.bfloat16), torch.empty(M, device=dev)
    ops.scalenorm_fwd(xs, g, y, rn)
    dyo, dxn = torch.randn(M, H, device=dev), torch.randn(M, H, device=dev).bfloat16()
    rs = torch.ones(H, device=dev)
    dx, dxb = torch.empty(M, H, device=dev), torch.empty(M, H, device=dev, dtype=torch.bfloat16)
    d_rs, d_g = torch.zeros(H, device=dev), torch.zeros(1, device=dev)
    for _ in range(2):
        ops.sublayer_bwd(dyo, dxn, xs, rn, g, rs, dx, dxb, d_rs, d_g)
    torch.cuda.synchronize()
else:
    from algonauts2025_b200.model import FmriEncoderConfig
    from algonauts2025_b200.pl_module import BrainModule
    from algonauts2025_b200.segment import SegmentData, synthetic_batch
    from algonauts2025_b200.trainer import MiniTrainer, default_optimizer

    torch.manual_seed(33)
    dims = {"text": (2, 3072), "audio": (2, 1024), "video": (2, 1408)}
    model = FmriEncoderConfig(n_subjects=4, modality_dropout=0.0).build(feature_dims=dims, n_outputs=1000, n_output_timesteps=100)
    module = BrainModule(model=model, loss=torch.nn.MSELoss(), optim_config=None, metrics={}, max_epochs=15)
    opt, sched = default_optimizer(model.parameters(), total_steps=100, model=model)
    tr = MiniTrainer(module, opt, sched)
    b = synthetic_batch(batch_size=16, seed=1)
    batch = SegmentData(data={k: v.cuda() for k, v in b.data.items()}, segments=b.segments)
    for _ in range(2):
        tr.train_step(batch)
    torch.cuda.synchronize()
print("done")
