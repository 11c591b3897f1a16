import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import algonauts2025_b200
from algonauts2025_b200 import ops
dev="cuda"
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
for O in (125, 250, 500, 1000):
    p, t = torch.randn(2560, O, 100, device=dev), torch.randn(2560, O, 100, device=dev)
    stats = torch.zeros(1, 6, O, device=dev, dtype=torch.float64)
    ops.pearson_stats(p, t, stats, layout="bdt"); torch.cuda.synchronize()
    ts=[]
    for _ in range(20):
        flush.zero_()
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); ops.pearson_stats(p, t, stats, layout="bdt"); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); ms=ts[10]
    print(f"cap={os.environ.get('TRIBE_PEARSON_MAX_CHUNKS','32')} parcels={O:5d} {ms*1e3:7.1f} us {8*p.numel()/ms/1e6:7.1f} GB/s", flush=True)
