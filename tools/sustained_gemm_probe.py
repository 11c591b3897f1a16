"""Sustained (power-capped) throughput of one GEMM shape repeated back to back for a few seconds: our 2-CTA tcgen05 kernel
vs cuBLAS (torch.matmul) on the same operands, with the SM clock sampled during each run.  Separates "how fast is the
kernel on an idle chip" (tools/gemm_bench.py) from "how many FLOPs per joule does it deliver at the board's power cap",
which is what bounds the train step.  Run under gpurun."""
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import algonauts2025_b200  # noqa: E402,F401
from algonauts2025_b200 import ops  # noqa: E402

dev = "cuda"


class Clock:
    def __init__(self):
        self.samples, self.stop = [], False

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append((float(out[0]), float(out[1])))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.2)


def sustained(fn, flops, seconds=3.0):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    clk = Clock()
    th = threading.Thread(target=clk.run, daemon=True)
    th.start()
    n, t0 = 0, time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(50):
            fn()
        n += 50
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    clk.stop = True
    th.join()
    ms = e0.elapsed_time(e1)
    s = clk.samples[len(clk.samples) // 3:] or [(0.0, 0.0)]
    mhz = sorted(x[0] for x in s)[len(s) // 2]
    watts = sorted(x[1] for x in s)[len(s) // 2]
    return flops * n / ms / 1e9, ms / n * 1e3, mhz, watts


def bf(*shape):
    return (torch.randn(*shape, device=dev) * 0.05).to(torch.bfloat16)


def main():
    M, H, F = 4768, 3072, 12288
    x, xF = bf(M, H), bf(M, F)
    w1, w2 = bf(F, H), bf(H, F)
    outF, outH = torch.empty(M, F, device=dev, dtype=torch.bfloat16), torch.empty(M, H, device=dev, dtype=torch.bfloat16)
    gW = torch.empty(F, H, device=dev)
    cases = [
        ("fwd ff1-shape  4768x12288x3072 (bf16 out, plain store)", lambda: ops.gemm(ops.kmajor(x), ops.kmajor(w1), outF, M, F, H, ldd=F),
         lambda: torch.matmul(x, w1.t(), out=outF), 2 * M * F * H),
        ("fwd ff2-shape  4768x3072x12288 (bf16 out, plain store)", lambda: ops.gemm(ops.kmajor(xF), ops.kmajor(w2), outH, M, H, F, ldd=H),
         lambda: torch.matmul(xF, w2.t(), out=outH), 2 * M * H * F),
        ("wgrad ff1-shape 12288x3072x4768 (fp32 out)", lambda: ops.gemm(ops.mnmajor(xF), ops.mnmajor(x), gW, F, H, M, ldd=H),
         lambda: torch.matmul(xF.t(), x, out=torch.empty(F, H, device=dev, dtype=torch.bfloat16)), 2 * M * F * H),
    ]
    for name, ours, cublas, flops in cases:
        for tag, fn in (("ours  ", ours), ("cuBLAS", cublas)):
            tf, us, mhz, watts = sustained(fn, flops)
            print(f"{name:58s} {tag}: {tf:7.1f} TFLOP/s sustained, {us:7.1f} us/launch, SM {mhz:6.0f} MHz, {watts:6.0f} W  "
                  f"-> {tf / max(mhz, 1.0) * 1000:6.1f} TFLOP/s per GHz, {tf / max(watts, 1.0):5.2f} TFLOP/J", flush=True)


if __name__ == "__main__":
    main()
