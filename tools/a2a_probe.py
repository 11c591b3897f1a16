import os, torch, torch.distributed as dist
rank=int(os.environ["RANK"]); ws=int(os.environ["WORLD_SIZE"]); local=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dist.init_process_group("nccl")
n,O,T=2560//ws,1000,100
x=torch.randn(n,O,T,device="cuda")
bounds=[(i*O//ws,(i+1)*O//ws) for i in range(ws)]
def once(tag):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e=[torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record()
    send=[x[:,a:b].contiguous() for a,b in bounds]
    e[1].record()
    recv=[torch.empty(n,bounds[rank][1]-bounds[rank][0],T,device="cuda") for _ in range(ws)]
    dist.all_to_all(recv,send)
    e[2].record()
    out=torch.cat(recv,0)
    e[3].record(); torch.cuda.synchronize()
    if rank==0: print(tag, "contiguous %.2f ms, all_to_all %.2f ms, cat %.2f ms"%(e[0].elapsed_time(e[1]),e[1].elapsed_time(e[2]),e[2].elapsed_time(e[3])), flush=True)
for i in range(4): once(f"list a2a #{i}")
# single-buffer variant
def once2(tag):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e=[torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    w=O//ws
    send=x.view(n,ws,w,T).permute(1,0,2,3).contiguous()   # (dst, n, w, T)
    e[1].record()
    recv=torch.empty(ws,n,w,T,device="cuda")
    dist.all_to_all_single(recv.view(-1),send.view(-1))
    e[2].record(); torch.cuda.synchronize()
    if rank==0: print(tag, "pack %.2f ms, all_to_all_single %.2f ms"%(e[0].elapsed_time(e[1]),e[1].elapsed_time(e[2])), flush=True)
for i in range(4): once2(f"single a2a #{i}")
dist.destroy_process_group()
