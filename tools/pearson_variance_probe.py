import os, sys, torch, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import algonauts2025_b200
from algonauts2025_b200 import ops
dev="cuda"
p, t = torch.randn(2560, 1000, 100, device=dev), torch.randn(2560, 1000, 100, device=dev)
stats = torch.zeros(1, 6, 1000, device=dev, dtype=torch.float64)
for _ in range(3): ops.pearson_stats(p, t, stats, layout="bdt")
torch.cuda.synchronize()
def run(n):
    ev=[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a,b in ev:
        stats.zero_(); a.record(); ops.pearson_stats(p, t, stats, layout="bdt"); b.record(); ops.pearson_finalize(stats[0])
    torch.cuda.synchronize()
    return [round(8*p.numel()/a.elapsed_time(b)/1e6) for a,b in ev]
print("idle GPU   GB/s per pass:", run(12))
print(subprocess.run(["nvidia-smi","--query-gpu=clocks.mem,clocks.max.mem,clocks.sm,temperature.gpu,power.draw","--format=csv,noheader"],capture_output=True,text=True).stdout.strip())
# heat the chip like the train legs do, then measure again
a=torch.randn(8192,8192,device=dev,dtype=torch.bfloat16); b=torch.randn(8192,8192,device=dev,dtype=torch.bfloat16)
import time
t0=time.time()
while time.time()-t0<20: (a@b); 
torch.cuda.synchronize()
print("after 20 s of GEMM load:", run(12))
print(subprocess.run(["nvidia-smi","--query-gpu=clocks.mem,clocks.max.mem,clocks.sm,temperature.gpu,power.draw","--format=csv,noheader"],capture_output=True,text=True).stdout.strip())
