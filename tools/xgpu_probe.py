"""torchrun probe of the NVLink step tail (parallel.ShardedStep / csrc/xgpu.cu): does symmetric memory come up, is there an
NVLS multicast mapping, what do the barrier and the fused reduce->Adam->multicast kernel cost on one encoder-layer
bucket (113 M parameters) as a function of the CTA count, multicast vs peer-pointer variant.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/xgpu_probe.py"""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import algonauts2025_b200  # noqa: E402
from algonauts2025_b200 import parallel  # noqa: E402
from algonauts2025_b200.model import FmriEncoder, FmriEncoderConfig  # noqa: E402
from algonauts2025_b200.trainer import default_optimizer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
algonauts2025_b200.load()


def log(*a):
    if rank == 0:
        print(*a, flush=True)


torch.manual_seed(3)
dims = {"text": (2, 3072), "audio": (2, 1024), "video": (2, 1408)}
depth = int(os.environ.get("PROBE_DEPTH", "2"))
model = FmriEncoder(dims, 1000, 100, FmriEncoderConfig(n_subjects=4), depth=depth)
opt, _ = default_optimizer(model.parameters(), total_steps=100, model=model)
t0 = time.perf_counter()
sh = parallel.ShardedStep(model, opt)
torch.cuda.synchronize()
log(f"symmetric memory up in {time.perf_counter() - t0:.2f} s: world {world}, multicast {sh.multicast}, "
    f"multicast_ptr grad {int(sh.handles['grad'].multicast_ptr):#x}, total {sh.flat.total / 1e6:.1f} M params")
opt.init_all_state()
fl = sh.flat
fl.grad.normal_()
torch.cuda.synchronize()
dist.barrier()

# barrier latency
for _ in range(5):
    sh.barrier(1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    sh.barrier(1)
e1.record()
torch.cuda.synchronize()
sh.check()
log(f"xgpu barrier: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per barrier kernel")

# one encoder-layer bucket
lo, hi = sh.owned[1][rank]
hyper = torch.tensor([0.9, 0.999, 1e-4, 1.0, 1e-8, 0.0, 0, 0], device="cuda")
n_own = hi - lo


def set_mode(mode):
    sh.mode = mode
    sh.out_multicast = sh.has_multicast and mode != "p2p"


for mode in (["staged", "nvls", "p2p"] if sh.has_multicast else ["staged", "p2p"]):
    set_mode(mode)
    for blocks in (148, 444, 888):
        sh.max_blocks = blocks
        for _ in range(2):
            sh.barrier(2)
            for x, y, b in sh.pieces(lo, hi):
                sh.launch(1, x, y, b, hyper.data_ptr())
        torch.cuda.synchronize()
        dist.barrier()
        e0.record()
        reps = 5
        for _ in range(reps):
            for x, y, b in sh.pieces(lo, hi):
                sh.launch(1, x, y, b, hyper.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        log(f"{mode:6s} blocks {blocks:4d}: {ms:7.3f} ms for {n_own / 1e6:.1f} M owned params  local HBM {(26 + (4 * world if mode == 'staged' else 0)) * n_own / ms / 1e6:7.1f} GB/s")
# copy-engine push of one bucket
for _ in range(2):
    sh.push_bucket(1)
torch.cuda.synchronize()
dist.barrier()
e0.record()
for _ in range(5):
    sh.push_bucket(1)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
sent = 4 * sum(b - a for r, (a, b) in enumerate(sh.owned[1]) if r != rank)
log(f"copy-engine push of one layer bucket: {ms:7.3f} ms for {sent / 1e6:.0f} MB -> {sent / ms / 1e6:.0f} GB/s out of this GPU")
sh.check()

# ---- pieces of the kernel, one at a time (which one is the ceiling?)
import ctypes  # noqa: E402
from algonauts2025_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
sink = torch.zeros(4, device="cuda")
big_lo, big_hi = [(x, y) for x, y, b in sh.pieces(lo, hi) if not b][0]
set_mode("nvls")
args = sh.kernel_args(1, big_lo, big_hi, False, hyper.data_ptr())
n_big = big_hi - big_lo
names = {0: "multimem.ld_reduce only", 1: "peer loads only", 2: "local p/m/v stream only", 3: "multimem.st bf16 only", 4: "ld_reduce, 4 groups/thread"}
for mode in (0, 4, 1, 2, 3):
    for blocks in (148, 592, 1184):
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for _ in range(2):
            lib.tribe_xgpu_probe(ctypes.byref(args), mode, blocks, ctypes.c_void_p(sink.data_ptr()), st)
        torch.cuda.synchronize()
        dist.barrier()
        e0.record()
        for _ in range(5):
            lib.tribe_xgpu_probe(ctypes.byref(args), mode, blocks, ctypes.c_void_p(sink.data_ptr()), st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        per = {0: 4, 4: 4, 1: 4 * world, 2: 24, 3: 2}[mode]
        log(f"piece [{names[mode]:28s}] blocks {blocks:4d}: {ms:7.3f} ms  {per * n_big / ms / 1e6:8.1f} GB/s ({per} B/param)")
sh.check()

# ---- interference with the persistent tcgen05 GEMM (FF1 shape): GEMM loop alone vs beside the step-tail kernel
M, N, K = 4768, 12288, 3072
xa = torch.randn(M, K, device="cuda").to(torch.bfloat16)
wb = torch.randn(N, K, device="cuda").to(torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
side = torch.cuda.Stream()


def gemm_loop(n):
    for _ in range(n):
        ops.gemm(ops.kmajor(xa), ops.kmajor(wb), out, M, N, K, ldd=N)


gemm_loop(5)
torch.cuda.synchronize()
dist.barrier()
e0.record()
gemm_loop(30)
e1.record()
torch.cuda.synchronize()
alone = e0.elapsed_time(e1) / 30
log(f"GEMM {M}x{N}x{K} alone: {alone * 1e3:.1f} us = {2 * M * N * K / alone / 1e9:.0f} TFLOP/s (carveout env {os.environ.get('TRIBE_XGPU_CARVEOUT', 'default=100')})")
for mc in (True, False):
    set_mode("nvls" if mc else "p2p")
    for blocks in (16, 148):
        sh.max_blocks = blocks
        torch.cuda.synchronize()
        dist.barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        side.wait_stream(torch.cuda.current_stream())
        e0.record()
        with torch.cuda.stream(side):
            s0.record()
            for _ in range(3):
                for x, y, b in sh.pieces(lo, hi):
                    sh.launch(1, x, y, b, hyper.data_ptr())
            s1.record()
        gemm_loop(30)
        e1.record()
        torch.cuda.synchronize()
        g = e0.elapsed_time(e1) / 30
        log(f"GEMM beside {'multicast' if mc else 'peer-ptr '} tail x3 blocks {blocks:4d}: GEMM {g * 1e3:7.1f} us ({g / alone:4.2f}x), "
            f"tail kernels {s0.elapsed_time(s1) / 3:6.3f} ms each, GEMM loop {g * 30:.2f} ms")
torch.cuda.synchronize()
dist.barrier()
side.wait_stream(torch.cuda.current_stream())
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
with torch.cuda.stream(side):
    s0.record()
    for _ in range(8):
        sh.push_bucket(1)
    s1.record()
gemm_loop(30)
e1.record()
torch.cuda.synchronize()
g = e0.elapsed_time(e1) / 30
log(f"GEMM beside 8 copy-engine bucket pushes: GEMM {g * 1e3:7.1f} us ({g / alone:4.2f}x), pushes {s0.elapsed_time(s1) / 8:6.3f} ms each")
sh.check()
# correctness of one launch of every mode against torch on the gathered gradients
for mode in (["staged", "nvls", "p2p"] if sh.has_multicast else ["staged", "p2p"]):
    set_mode(mode)
    sh.max_blocks = 0
    fl.grad.normal_()
    fl.adam_m.zero_(), fl.adam_v.zero_()
    torch.cuda.synchronize()
    dist.barrier()
    allg = [torch.empty(fl.total, device="cuda") for _ in range(world)]
    dist.all_gather(allg, fl.grad.clone())
    gmean = sum(a[lo:hi] for a in allg) / world
    p_before = fl.flat[lo:hi].clone()
    if mode == "staged":
        sh.push_bucket(1)
    sh.barrier(3)
    hyper_eps1 = torch.tensor([0.9, 0.999, 1e-4, 1.0, 1.0, 0.0, 0, 0], device="cuda")  # eps = 1: the update depends on the gradient SCALE (mean, not sum)
    for x, y, b in sh.pieces(lo, hi):
        sh.launch(1, x, y, b, hyper_eps1.data_ptr())
    sh.barrier(4)
    torch.cuda.synchronize()
    m = gmean * (1 - 0.9)
    v = (1 - 0.999) * gmean * gmean
    want = p_before - 1e-4 * (m / (v.sqrt() * 1.0 + 1.0))
    err = float((fl.flat[lo:hi] - want).abs().max())
    shadow_all = [torch.empty(fl.total, device="cuda", dtype=torch.bfloat16) for _ in range(world)]
    dist.all_gather(shadow_all, fl.bf16.clone())
    ok_shadow = all(torch.equal(s_, shadow_all[0]) for s_ in shadow_all)
    own_ok = torch.equal(fl.bf16[lo:hi], fl.flat[lo:hi].to(torch.bfloat16))
    print(f"rank {rank} mode {mode}: adam max err {err:.2e}; shadow equal on all ranks {ok_shadow}; shadow == bf16(master) on owned slice {own_ok}", flush=True)
    sh.check()
dist.barrier()
os._exit(0)
