"""Per-shape throughput of the tcgen05 GEMM on the TRIBE train-step shapes (run under gpurun).
Each launch is timed with CUDA events after an L2 flush; prints TFLOP/s vs the measured bf16 peak."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import algonauts2025_b200  # noqa: E402
from algonauts2025_b200 import ops  # noqa: E402

M, H, F, B, T, HEADS = 4768, 3072, 12288, 16, 298, 8
DH, TP = H // HEADS, 304
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)


def bf(*shape):
    return (torch.randn(*shape, device=dev) * 0.05).to(torch.bfloat16)


def timeit(fn, flops, iters=8):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    med = ts[len(ts) // 2]
    return med, flops / med / 1e9


def main():
    block_ns = [int(x) for x in os.environ.get("BLOCK_NS", "0").split(",")]
    x, xF = bf(M, H), bf(M, F)
    w_qkv, w_o, w1, w2 = bf(3 * H, H), bf(H, H), bf(F, H), bf(H, F)
    out3, outH, outF = torch.empty(M, 3 * H, device=dev, dtype=torch.bfloat16), torch.empty(M, H, device=dev, dtype=torch.bfloat16), torch.empty(M, F, device=dev, dtype=torch.bfloat16)
    outHf = torch.empty(M, H, device=dev)
    gW1, gW2, gQKV, gO = torch.empty(F, H, device=dev), torch.empty(H, F, device=dev), torch.empty(3 * H, H, device=dev), torch.empty(H, H, device=dev)
    dy3 = bf(M, 3 * H)
    qkv = bf(M, 3 * H)
    S = torch.empty(B * HEADS, T, TP, device=dev)
    P = bf(B * HEADS, T, TP)
    res = torch.randn(M, H, device=dev)
    cases = {}
    for bn in block_ns:
        tag = f"bn{bn}"
        cases[f"fwd qkv   M{M} N{3*H} K{H} {tag}"] = (lambda bn=bn: ops.gemm(ops.kmajor(x), ops.kmajor(w_qkv), out3, M, 3 * H, H, ldd=3 * H, block_n=bn), 2 * M * 3 * H * H)
        cases[f"fwd out   M{M} N{H} K{H} {tag} (+res f32)"] = (lambda bn=bn: ops.gemm(ops.kmajor(x), ops.kmajor(w_o), outHf, M, H, H, ldd=H, epilogue=ops.EPI_RESIDUAL, res=res, ld_res=H, block_n=bn), 2 * M * H * H)
        cases[f"fwd ff1   M{M} N{F} K{H} {tag} (+gelu)"] = (lambda bn=bn: ops.gemm(ops.kmajor(x), ops.kmajor(w1), outF, M, F, H, ldd=F, epilogue=ops.EPI_GELU, aux_out=xF, ld_aux=F, block_n=bn), 2 * M * F * H)
        cases[f"fwd ff2   M{M} N{H} K{F} {tag} (+res f32)"] = (lambda bn=bn: ops.gemm(ops.kmajor(xF), ops.kmajor(w2), outHf, M, H, F, ldd=H, epilogue=ops.EPI_RESIDUAL, res=res, ld_res=H, block_n=bn), 2 * M * H * F)
        cases[f"dgrad ff2 M{M} N{F} K{H} {tag} (Bmn, gelu')"] = (lambda bn=bn: ops.gemm(ops.kmajor(x), ops.mnmajor(w2), outF, M, F, H, ldd=F, epilogue=ops.EPI_GELU_BWD, aux_in=xF, ld_aux=F, block_n=bn), 2 * M * F * H)
        cases[f"dgrad ff1 M{M} N{H} K{F} {tag} (Bmn)"] = (lambda bn=bn: ops.gemm(ops.kmajor(xF), ops.mnmajor(w1), outH, M, H, F, ldd=H, block_n=bn), 2 * M * H * F)
        cases[f"dgrad qkv M{M} N{H} K{3*H} {tag} (Bmn)"] = (lambda bn=bn: ops.gemm(ops.kmajor(dy3), ops.mnmajor(w_qkv), outH, M, H, 3 * H, ldd=H, block_n=bn), 2 * M * H * 3 * H)
        cases[f"wgrad ff1 M{F} N{H} K{M} {tag} (AmnBmn f32)"] = (lambda bn=bn: ops.gemm(ops.mnmajor(xF), ops.mnmajor(x), gW1, F, H, M, ldd=H, block_n=bn), 2 * M * F * H)
        cases[f"wgrad ff2 M{H} N{F} K{M} {tag} (AmnBmn f32)"] = (lambda bn=bn: ops.gemm(ops.mnmajor(x), ops.mnmajor(xF), gW2, H, F, M, ldd=F, block_n=bn), 2 * M * F * H)
        cases[f"wgrad qkv M{3*H} N{H} K{M} {tag} (AmnBmn f32)"] = (lambda bn=bn: ops.gemm(ops.mnmajor(dy3), ops.mnmajor(x), gQKV, 3 * H, H, M, ldd=H, block_n=bn), 2 * M * 3 * H * H)
        cases[f"wgrad out M{H} N{H} K{M} {tag} (AmnBmn f32)"] = (lambda bn=bn: ops.gemm(ops.mnmajor(x), ops.mnmajor(x), gO, H, H, M, ldd=H, block_n=bn), 2 * M * H * H)
    q_op = ops.Operand(qkv, inner=3 * H, rows=T, row_stride=3 * H, batch=B, batch_stride=T * 3 * H, zin_stride=DH, zdiv=HEADS)
    k_op = ops.Operand(qkv, inner=3 * H, rows=T, row_stride=3 * H, batch=B, batch_stride=T * 3 * H, inner_off=H, zin_stride=DH, zdiv=HEADS)
    cases["attn S=QK^T batched 128x(298x298x384)"] = (lambda: ops.gemm(q_op, k_op, S, T, T, DH, ldd=TP, batch=B * HEADS, z_inner=HEADS, d_zo=HEADS * T * TP, d_zi=T * TP, alpha=0.05), 2 * B * HEADS * T * T * DH)
    p_op = ops.Operand(P, inner=TP, rows=T, row_stride=TP, batch=B * HEADS, batch_stride=T * TP)
    v_op = ops.Operand(qkv, inner=3 * H, rows=T, row_stride=3 * H, batch=B, batch_stride=T * 3 * H, mn_major=True, inner_off=2 * H, zin_stride=DH, zdiv=HEADS)
    cases["attn O=PV   batched 128x(298x384x304)"] = (lambda: ops.gemm(p_op, v_op, outH, T, DH, TP, ldd=H, batch=B * HEADS, z_inner=HEADS, d_zo=T * H, d_zi=DH), 2 * B * HEADS * T * T * DH)
    peak = 1383.5
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]
    # cuBLAS reference point (library GEMM, for context only)
    a, b = bf(M, H), bf(F, H)
    ms, tf = timeit(lambda: torch.matmul(a, b.t()), 2 * M * F * H)
    print(f"{'cuBLAS bf16 M4768 N12288 K3072 (context)':58s} {ms*1e3:8.1f} us {tf:8.1f} TFLOP/s {tf/peak:6.1%} of burst peak")
    for name, (fn, flops) in cases.items():
        ops.SPLITK = True
        ms, tf = timeit(fn, flops)
        ops.SPLITK = False
        ms0, tf0 = timeit(fn, flops)
        ops.SPLITK = True
        print(f"{name:58s} {ms*1e3:8.1f} us {tf:8.1f} TFLOP/s {tf/peak:6.1%} of burst peak | no split-K tail: {ms0*1e3:8.1f} us {tf0:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
