"""Achieved HBM bandwidth of the bandwidth-bound kernels at the TRIBE shapes (run under gpurun).
ALGORITHMIC bytes / CUDA-event time, L2 flushed between launches; peak = MEASURED_PEAKS.json hbm_gbs.
The flush READS a 256 MB buffer (BW_FLUSH=read, default): L2 ends up full of clean lines of foreign data.  The round-1
flush (BW_FLUSH=write: zero-fill) left 126 MB of DIRTY lines whose write-back then ran concurrently with the measured
kernel — for kernels that move only 100-250 MB that added up to +50 % HBM traffic to their clock.
Also prints the Pearson-eval metric of BASELINE.json: parcel-TRs/s on (N_TR = 256 000, 1000) fp32 matrices."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import algonauts2025_b200  # noqa: E402
from algonauts2025_b200 import ops  # noqa: E402

dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)
PEAK = 6547.2
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
RESULTS = {}


FLUSH_MODE = os.environ.get("BW_FLUSH", "read")


def timeit(name, fn, nbytes, iters=10, extra=""):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if FLUSH_MODE == "write":
            flush.zero_()
        else:
            flush.view(torch.int32).sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    gbs = nbytes / ms / 1e6
    RESULTS[name] = {"us": ms * 1e3, "GBps": gbs, "frac": gbs / PEAK}
    print(f"{name:52s} {ms*1e3:9.1f} us {gbs:8.1f} GB/s {gbs/PEAK:6.1%} of measured HBM peak {extra}", flush=True)
    return ms


def main():
    B, T, TQ, H, F, O = 16, 298, 100, 3072, 12288, 1000
    M = B * T
    # Pearson eval: config 5 (N_TR = 256 000 rows x 1000 parcels, fp32 pred + target resident on the device)
    n_tr = 256_000
    pred, true = torch.randn(n_tr, O, device=dev), torch.randn(n_tr, O, device=dev)
    stats = torch.zeros(1, 6, O, device=dev, dtype=torch.float64)
    ms = timeit("pearson_stats row-major (256000 x 1000)", lambda: ops.pearson_stats(pred, true, stats, layout="no"), 8 * n_tr * O)
    print(f"    Pearson eval: {n_tr * O / ms / 1e6:.1f} G parcel-TRs/s (roofline {PEAK / 8:.0f} G/s)")
    RESULTS["pearson_parcel_trs_per_s"] = n_tr * O / ms * 1e3
    del pred, true
    p3, t3 = torch.randn(64, O, TQ, device=dev), torch.randn(64, O, TQ, device=dev)
    stats.zero_()
    timeit("pearson_stats (b,d,t) layout (64 x 1000 x 100)", lambda: ops.pearson_stats(p3, t3, stats, layout="bdt"), 8 * p3.numel())
    pr, tg = torch.randn(B, O, TQ, device=dev), torch.randn(B, O, TQ, device=dev)
    timeit("mse fwd+grad (16 x 1000 x 100)", lambda: ops.mse_fwd_bwd(pr, tg), 12 * pr.numel())
    big_p, big_t = torch.randn(64 * 16, O, TQ, device=dev), torch.randn(64 * 16, O, TQ, device=dev)
    timeit("mse fwd+grad (1024 x 1000 x 100)", lambda: ops.mse_fwd_bwd(big_p, big_t), 12 * big_p.numel())
    del big_p, big_t
    # adaptive average pool (B, O, T) -> (B, O, T'): 1.592 MB / window
    x = torch.randn(B, O, T, device=dev)
    timeit("adaptive_avg_pool fwd (16 x 1000 x 298 -> 100)", lambda: ops.adaptive_avg_pool_fwd(x, TQ), 4 * B * O * (T + TQ))
    xb = torch.randn(256, O, T, device=dev)
    timeit("adaptive_avg_pool fwd (256 x 1000 x 298 -> 100)", lambda: ops.adaptive_avg_pool_fwd(xb, TQ), 4 * 256 * O * (T + TQ))
    dy = torch.randn(256, O, TQ, device=dev)
    timeit("adaptive_avg_pool bwd (256 x 1000 x 100 -> 298)", lambda: ops.adaptive_avg_pool_bwd(dy, T), 4 * 256 * O * (T + TQ))
    del xb, dy
    # ingest: 19.7 MB / window
    for name, (L, D) in {"text": (2, 3072), "audio": (2, 1024), "video": (2, 1408)}.items():
        feats = torch.randn(B, L, D, T, device=dev)
        out = torch.empty(M, L * D, device=dev, dtype=torch.bfloat16)
        timeit(f"ingest {name} (16 x {L} x {D} x 298) fp32 -> bf16", lambda: ops.ingest_features(feats, out, 0, False), 6 * feats.numel())
    xs = torch.randn(M, H, device=dev)
    g = torch.ones(1, device=dev)
    y, rn = torch.empty(M, H, device=dev, dtype=torch.bfloat16), torch.empty(M, device=dev)
    timeit("scalenorm fwd (4768 x 3072)", lambda: ops.scalenorm_fwd(xs, g, y, rn), 6 * xs.numel())
    dyo, dxn = torch.randn(M, H, device=dev), torch.randn(M, H, device=dev).bfloat16()
    rs = torch.ones(H, device=dev)
    dx, dxb = torch.empty(M, H, device=dev), torch.empty(M, H, device=dev, dtype=torch.bfloat16)
    d_rs, d_g = torch.zeros(H, device=dev), torch.zeros(1, device=dev)
    timeit("sublayer_bwd tail (4768 x 3072)", lambda: ops.sublayer_bwd(dyo, dxn, xs, rn, g, rs, dx, dxb, d_rs, d_g), 16 * xs.numel())
    S = torch.randn(128, T, 304, device=dev)
    P = torch.empty(128, T, 304, device=dev, dtype=torch.bfloat16)
    timeit("softmax fwd (128 x 298 x 304)", lambda: ops.softmax_fwd(S, P, T), 6 * S.numel())
    dS = torch.empty_like(P)
    timeit("softmax bwd (128 x 298 x 304)", lambda: ops.softmax_bwd(P, S, dS, 0.05, T), 8 * S.numel())
    hb = torch.randn(M, F, device=dev).bfloat16()
    outc = torch.zeros(F, device=dev)
    timeit("colsum bf16 (4768 x 12288)", lambda: ops.colsum(hb, outc), 2 * hb.numel())
    w = torch.randn(200_000_000, device=dev)
    wb = torch.empty(200_000_000, device=dev, dtype=torch.bfloat16)
    timeit("cast fp32 -> bf16 (200 M params)", lambda: ops.cast_f32_bf16(w, wb), 6 * w.numel())
    xt = torch.randn(B, T, H, device=dev).bfloat16()
    yp = torch.empty(B, TQ, H, device=dev, dtype=torch.bfloat16)
    timeit("token pool fwd (16 x 298 x 3072 -> 100)", lambda: ops.token_pool_fwd(xt, yp, B, T, TQ, H), 2 * B * H * (T + TQ))
    del w, wb
    # ---- SURVEY §8f rows
    big = torch.randn(2560, O, TQ, device=dev)
    bigt = torch.randn(2560, O, TQ, device=dev)
    stats.zero_()
    timeit("pearson_stats (b,d,t) layout (2560 x 1000 x 100)", lambda: ops.pearson_stats(big, bigt, stats, layout="bdt"), 8 * big.numel())
    _, coef = ops.pearson_loss_fwd(big, bigt, layout="bdt")
    up = torch.ones(1, device=dev)
    timeit("pearson_loss bwd (2560 x 1000 x 100)", lambda: ops.pearson_loss_bwd(big, bigt, coef, up, layout="bdt"), 12 * big.numel())
    timeit("smooth_l1 fwd+grad (2560 x 1000 x 100)", lambda: ops.point_loss_fwd_bwd(big, bigt, ops.LOSS_SMOOTH_L1, 1.0), 12 * big.numel())
    timeit("pearson_loss fwd (16 x 1000 x 100) [train batch]", lambda: ops.pearson_loss_fwd(pr, tg, layout="bdt"), 8 * pr.numel())
    del big, bigt
    members = torch.randn(8, 25_600, O, device=dev)
    wts = ops.ensemble_weights(torch.rand(8, O, device=dev), 0.3)
    timeit("ensemble average (8 members x 25600 x 1000)", lambda: ops.ensemble_average(members, wts), 4 * members.numel() + 4 * 25_600 * O)
    del members
    from algonauts2025_b200 import windows as W
    store = W.TimelineStore()
    for tl in range(4):
        store.add("text", f"t{tl}", torch.randn(2, 3072, 1500), start=0.0, frequency=2.0)
    wins = [(f"t{i % 4}", -4.47 + 149.0 * (i // 4)) for i in range(16)]
    timeit("window gather text (16 x 2 x 3072 x 298)", lambda: store.assemble("text", wins), 8 * 16 * 2 * 3072 * 298)
    avg, par = torch.zeros(200_000_000, device=dev), torch.randn(200_000_000, device=dev)
    timeit("swa update (200 M params)", lambda: ops.swa_update(avg, par, 3), 12 * avg.numel())
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(RESULTS, open(os.path.join(ROOT, "gpurun_out", "bw_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
