"""1-GPU probe: can a small CTA share an SM with the persistent tcgen05 GEMM CTAs (320 threads x 168 registers, ~200 KB of
dynamic shared memory, 2-CTA clusters, static tile schedule)?  A spin kernel occupies k CTAs for a few ms on a side
stream while a loop of GEMMs runs on the main stream; if the GEMMs slow down by ~1.7x the spin CTAs took whole SMs away
(no co-residency) and the static schedule pays a second wave."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import algonauts2025_b200  # noqa: E402
from algonauts2025_b200 import _lib, ops  # noqa: E402

lib = algonauts2025_b200.load()
torch.cuda.set_device(0)
sink = torch.zeros(4, device="cuda", dtype=torch.int32)
side = torch.cuda.Stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
shapes = {"ff1 4768x12288x3072 (2-CTA kernel)": (4768, 12288, 3072), "out 4768x3072x3072 (2-CTA kernel)": (4768, 3072, 3072)}
for name, (M, N, K) in shapes.items():
    xa = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    wb = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)

    def loop(n):
        for _ in range(n):
            ops.gemm(ops.kmajor(xa), ops.kmajor(wb), out, M, N, K, ldd=N)

    loop(5)
    torch.cuda.synchronize()
    e0.record()
    loop(20)
    e1.record()
    torch.cuda.synchronize()
    alone = e0.elapsed_time(e1) / 20
    print(f"{name}: alone {alone * 1e3:.1f} us", flush=True)
    for sms in (0, 148):
        if sms:
            ops.gemm_set_sm_limit(sms)
        for blocks, threads, carve in ((1, 32, -1), (1, 32, 100), (16, 32, 100), (16, 128, 100), (148, 32, 100), (148, 128, 100), (16, 1024, 100)):
            torch.cuda.synchronize()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                lib.tribe_debug_spin(blocks, threads, 0.012, carve, ctypes.c_void_p(sink.data_ptr()), ctypes.c_void_p(side.cuda_stream))
            e0.record()
            loop(20)
            e1.record()
            torch.cuda.synchronize()
            g = e0.elapsed_time(e1) / 20
            print(f"   gemm-sm-limit {sms or 'off'}: beside spin {blocks:3d} CTAs x {threads:4d} thr (carveout {carve:3d}): {g * 1e3:7.1f} us ({g / alone:4.2f}x)", flush=True)
        ops.gemm_set_sm_limit(0)
    # reverse order: GEMMs already running when a small kernel arrives — how long until it gets an SM?
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    loop(3)
    side.wait_stream(torch.cuda.current_stream())
    loop(20)
    with torch.cuda.stream(side):
        s0.record()
        lib.tribe_debug_spin(1, 32, 1e-6, 100, ctypes.c_void_p(sink.data_ptr()), ctypes.c_void_p(side.cuda_stream))
        s1.record()
    torch.cuda.synchronize()
    print(f"   a 1-CTA kernel launched behind 20 queued GEMMs took {s0.elapsed_time(s1) * 1e3:.1f} us to finish (GEMM {alone * 1e3:.0f} us each)", flush=True)
