"""bench.py — TRIBE FmriEncoder train step on B200 (BASELINE.json metric: train windows/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--contrastive 0|1] [--batch 16]

One "step" = one Lightning-style automatic-optimisation step of ``BrainModule`` (zero_grad -> training_step: forward,
MSE loss [-> contrastive branch] -> backward -> Adam -> OneCycleLR) on one batch of 16 synthetic windows per GPU of the
full TRIBE shape (text 2x3072, audio 2x1024, video 2x1408 feature stacks at T=298, 1000 parcels x 100 TRs, 4 subjects).
`value`  : windows/s with the inputs already resident in HBM (CUDA events, max over ranks).
`e2e`    : the same metric through the public API from pinned HOST batches (H2D inside the timed region, loss read back).
`roofline`: all tcgen05 GEMM launches of the timed steps, algorithmic FLOPs / summed CUDA-event time, against the
            measured sustained bf16 peak (MEASURED_PEAKS.json).
`cpu_baseline` / `--impl reference`: the CPU oracle port of the reference path (oracle/) on the host cores.
Under torchrun (N > 1) every rank trains its own 16 windows (weak scaling); gradients travel by copy engine over NVLink
into their owner's staging buffer during the backward pass and one fused kernel per owned range reduces them, runs the
rank-sharded Adam and multicasts the bf16 shadow weights (parallel.ShardedStep, csrc/xgpu.cu) — no NCCL in the step.
Extra keys of the same JSON line (each with its own timing, BASELINE.json configs 4 and 5 and the reference's default
recipe): `contrastive_on`, `ensemble` (N > 1), `eval_predict`, `eval_sweep`, `pearson_eval`.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FEATURE_DIMS = {"text": (2, 3072), "audio": (2, 1024), "video": (2, 1408)}
WORKLOAD = "TRIBE FmriEncoder train step: 4 subjects, 16 windows/GPU x (text 2x3072 + audio 2x1024 + video 2x1408) x T=298 -> 1000 parcels x 100 TRs"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"tflops": float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1590.0))), "hbm": float(p.get("hbm_gbs", 6650.0)), "src": "measured"}
    return {"tflops": 1590.0, "hbm": 6650.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float | None = None, t1: float | None = None):
        """Median SM clock and throttle reasons of the samples that arrived inside [t0, t1] (perf_counter; the timed
        region), or of every sample taken under load when the window caught none."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        inside = [ln for t, ln in self.lines if t0 is None or (t0 <= t <= (t1 if t1 is not None else t))]
        window = "timed region"
        if not inside:
            inside, window = [ln for _, ln in self.lines], "whole loaded run (timed region shorter than the sampling period)"
        sm, mx, reasons = [], None, set()
        for ln in inside:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])), (mx := float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm), "window": window}


def algorithmic_flops_per_window(contrastive: bool) -> float:
    """forward 557.3 GFLOP/window, train step = 3x; the contrastive branch adds a second projector +
    encoder pass, the head GEMM and the InfoNCE logits."""
    t, h, depth, n_out, k_proj = 298, 3072, 8, 1000, 11008
    proj = 2 * t * k_proj * (h // 3)
    enc = depth * (4 * 2 * t * h * h + 2 * 2 * t * h * 4 * h)
    attn = depth * 2 * (2 * t * t * h)
    readout = 2 * t * h * n_out
    fwd = proj + enc + attn + readout
    if not contrastive:
        return 3.0 * fwd
    return 3.0 * (fwd + proj + enc + attn + 2 * t * 2816 * h + 2 * (16 * t) * t * h)


# ------------------------------------------------------------------------------------------------------ CPU (oracle) arm
REF_WINDOWS_PER_STEP = 4  # bounded sample of the 16-window step: enough rows (1192) for the host BLAS to run efficiently

def cpu_reference_steps(steps: int, warmup: int, windows_per_step: int, contrastive: bool):
    """The reference's own CPU path (PyTorch fp32): oracle port of model.py / pl_module.py / x_transformers, stock
    torch.optim.Adam, all host threads."""
    from oracle import tribe_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(33)
    cfg = O.OracleConfig(n_subjects=4, modality_dropout=0.3, contrastive_enabled=contrastive)
    model = O.OracleFmriEncoder(FEATURE_DIMS, 1000, 100, cfg)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    batch = O.synthetic_batch(batch_size=windows_per_step, seed=1234)
    model.train()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss, *_ = O.run_step(model, batch)
        loss.backward()
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"windows_per_s": windows_per_step * steps / total, "ms_per_step": 1e3 * total / steps, "cores": torch.get_num_threads(),
            "sample": f"{steps} train step(s) of {windows_per_step} window(s) each after {warmup} warm-up, fp32, torch CPU, Adam"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_steps(args.steps, args.warmup, REF_WINDOWS_PER_STEP, bool(args.contrastive))
    line = {"impl": "reference", "metric": "train windows/s", "value": r["windows_per_s"], "unit": "windows/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "windows_per_step": REF_WINDOWS_PER_STEP, "contrastive": bool(args.contrastive)},
            "cpu_baseline": {"value": r["windows_per_s"], "unit": "windows/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["windows_per_s"], "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    pe = cpu_pearson_baseline()
    line["pearson_eval"] = {"metric": "Pearson eval parcel-TRs/s", "value": pe["value"], "unit": pe["unit"], "cpu_baseline": pe}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------ Pearson eval leg
EVAL_WINDOWS, EVAL_PARCELS, EVAL_TRS = 2560, 1000, 100  # Movie10-shaped held-out set, 256 M parcel-TRs


def cpu_pearson_baseline(n_windows: int = 256):
    """The reference's literal evaluation loop (main.py:470-476): host rearrange + 1000 x scipy.stats.pearsonr."""
    import numpy as np
    from oracle import tribe_oracle as O

    rng = np.random.default_rng(7)
    trues = rng.standard_normal((n_windows, EVAL_PARCELS, EVAL_TRS), dtype=np.float32)
    preds = (0.2 * trues + rng.standard_normal(trues.shape, dtype=np.float32)).astype(np.float32)
    t0 = time.perf_counter()
    r = O.multidim_pearson_scipy(preds, trues)
    dt = time.perf_counter() - t0
    assert r.shape == (EVAL_PARCELS,)
    n = n_windows * EVAL_PARCELS * EVAL_TRS
    return {"value": n / dt, "unit": "parcel-TRs/s", "cores": 1, "kind": "port",
            "sample": f"{n_windows} windows x {EVAL_PARCELS} parcels x {EVAL_TRS} TRs through the main.py:470-476 scipy.stats.pearsonr loop"}


def eval_predict_leg(model, rank: int, world: int, steps: int, batch: int = 64):
    """BASELINE.json config 5, first half: eval-mode batched predict (windows sharded across ranks, no collective) feeding
    the per-parcel Pearson statistics — ``metrics.compute_multidim_pearson`` (main.py:459-477) from pinned host batches."""
    import torch.distributed as dist

    from algonauts2025_b200 import metrics
    from algonauts2025_b200.segment import DevicePrefetcher, synthetic_batch

    host = [synthetic_batch(batch_size=batch, seed=4321 + 31 * rank + i, pin=True) for i in range(2)]
    prefetch = DevicePrefetcher(())
    dev = prefetch.resident(host)
    model.eval()
    with torch.no_grad():
        for i in range(2):
            model(dev[i % 2])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            y = model(dev[i % 2])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        t0 = time.perf_counter()
        r = metrics.compute_multidim_pearson(model, prefetch.feed(host[j % 2] for j in range(steps)))
        e2e_s = time.perf_counter() - t0
    assert r.shape == (1000,) and y.shape == (batch, 1000, 100)
    t = torch.tensor([ms, e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_s = (float(x) for x in t.cpu())
    model.train()
    h2d = sum(v.numel() * v.element_size() for v in host[0].data.values())
    flops = algorithmic_flops_per_window(False) / 3.0
    return {"metric": "eval predict windows/s", "value": world * batch * steps / (ms / 1e3), "unit": "windows/s", "ms_per_batch": ms / steps,
            "batch_per_gpu": batch, "tflops": flops * batch * steps * world / (ms / 1e3) / 1e12 / world,
            "e2e": {"value": world * batch * steps / e2e_s, "unit": "windows/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4000,
                    "api": "metrics.compute_multidim_pearson(model, DevicePrefetcher(pinned host batches)) -> r[1000] on host"}}


def pearson_eval_leg(rank: int, world: int, steps: int, warmup: int, peaks, with_cpu: bool):
    """Per-parcel Pearson r over (windows, parcels, TRs) prediction/target tensors, parcels sharded 1000/G per rank
    (no data-path collective: shards are independent; r is gathered once at the end, outside the kernel timing)."""
    import torch.distributed as dist

    from algonauts2025_b200 import metrics, ops, parallel

    lo, hi = parallel.parcel_bounds(EVAL_PARCELS, world)[rank]
    g = torch.Generator(device="cuda").manual_seed(99 + rank)
    trues = torch.randn(EVAL_WINDOWS, hi - lo, EVAL_TRS, device="cuda", generator=g)
    preds = 0.2 * trues + torch.randn(EVAL_WINDOWS, hi - lo, EVAL_TRS, device="cuda", generator=g)
    stats = torch.zeros(1, 6, hi - lo, device="cuda", dtype=torch.float64)
    shift = torch.empty(2, hi - lo, device="cuda", dtype=torch.float32)

    def one():
        ops.zero_(stats)
        ops.pearson_pick_shift(preds, trues, shift, layout="bdt")  # per-parcel pivots (first row): centred moments
        ops.pearson_stats(preds, trues, stats, layout="bdt", shift=shift)
        return ops.pearson_finalize(stats[0])[0]

    for _ in range(max(warmup, 3)):
        r = one()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # (1) the whole pass (zero -> pivots -> statistics -> finalize: four launches) replayed as ONE CUDA graph: at 125 parcels
    # per rank the statistics kernel runs ~45 us and launching the pass from Python would cost more than running it
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, capture_error_mode="thread_local"):
        r = one()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(steps):  # operands (2 x 1.02 GB / G) exceed L2, every pass streams from HBM
        graph.replay()
    g1.record()
    torch.cuda.synchronize()
    total_ms = g0.elapsed_time(g1)
    # (2) the statistics kernel alone (the roofline's kernel), eager, CUDA events around each launch
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for b, c in ev:
        ops.zero_(stats)
        ops.pearson_pick_shift(preds, trues, shift, layout="bdt")
        b.record()
        ops.pearson_stats(preds, trues, stats, layout="bdt", shift=shift)
        c.record()
        ops.pearson_finalize(stats[0])
    torch.cuda.synchronize()
    kern_ms = sum(b.elapsed_time(c) for b, c in ev) / steps
    # end to end from HOST arrays (what main.py:470-473 holds after the predict loop): chunked H2D + statistics + r back
    n_e2e = 640 // world
    hp = preds[:n_e2e].cpu().pin_memory()
    ht = trues[:n_e2e].cpu().pin_memory()
    metrics.pearson_from_host(hp, ht)  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r_host = metrics.pearson_from_host(hp, ht)
    e2e_s = time.perf_counter() - t0
    assert r_host.shape == (hi - lo,)
    t = torch.tensor([total_ms, kern_ms, e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kern_ms, e2e_s = (float(x) for x in t.cpu())
    r_all = parallel.gather_parcels(r, EVAL_PARCELS)
    if rank != 0:
        return None
    n_total = EVAL_WINDOWS * EVAL_PARCELS * EVAL_TRS
    shard_bytes = 8.0 * EVAL_WINDOWS * (hi - lo) * EVAL_TRS
    gbs = shard_bytes / (kern_ms / 1e3) / 1e9
    out = {"metric": "Pearson eval parcel-TRs/s", "value": n_total * steps / (total_ms / 1e3), "unit": "parcel-TRs/s", "n_gpus": world,
           "scaling": "strong", "ms_per_pass": total_ms / steps, "mean_r": float(r_all.mean()),
           "config": {"workload": f"per-parcel Pearson r, {EVAL_WINDOWS} windows x {EVAL_PARCELS} parcels x {EVAL_TRS} TRs fp32 (b,d,t) layout, parcels sharded {EVAL_PARCELS}/{world} per GPU",
                      "l2": "operands larger than L2", "launch": "the four launches of a pass replayed as one CUDA graph"},
           "e2e": {"value": n_e2e * EVAL_PARCELS * EVAL_TRS / e2e_s,
                   "unit": "parcel-TRs/s", "h2d_bytes_per_step": 8 * n_e2e * (hi - lo) * EVAL_TRS, "d2h_bytes_per_step": 4 * (hi - lo),
                   "sample": f"{n_e2e} windows per rank from pinned host arrays via metrics.pearson_from_host"},
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                        # dram__bytes_read+write of one launch at this exact shape on 1 GPU from an ncu --set full capture, with its
                        # provenance (profiles/ncu_reference.json); NOT measured by this run
                        "traffic": ncu_reference().get("pearson_bdt_dram_bytes", {}).get("value") if world == 1 else None,
                        "traffic_source": ncu_reference().get("pearson_bdt_dram_bytes", {}).get("source"), "algorithmic_bytes": shard_bytes,
                        "kernel": "pearson_bdt_kernel<4>", "peak_source": peaks["src"] + " copy bandwidth", "bytes_per_parcel_tr": 8}}
    if with_cpu:
        out["cpu_baseline"] = cpu_pearson_baseline()
    return out


# ------------------------------------------------------------------------------------------------------ GPU arm
def ncu_reference():
    """ncu-derived context numbers with their provenance (profiles/ncu_reference.json); never measured by this run."""
    path = os.path.join(ROOT, "profiles", "ncu_reference.json")
    try:
        return json.load(open(path))
    except Exception:  # noqa: BLE001
        return {}


def _barrier(world):
    import torch.distributed as dist

    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(values, world):
    import torch.distributed as dist

    t = torch.tensor(values, device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def train_leg(args, rank, world, local, *, contrastive: bool, dp: str, seed: int, steps: int, warmup: int, roofline: bool, e2e: bool,
              clocks: ClockSampler | None = None, keep_model: bool = False):
    """One training configuration: build the model (+ optimizer, data-parallel step tail, whole-step graphs), W warm-up
    steps, K timed steps on device-resident batches, K steps end to end from pinned host batches, and (``roofline``) K
    instrumented steps with a CUDA-event pair around every tcgen05 GEMM launch."""
    import algonauts2025_b200
    from algonauts2025_b200 import ops, parallel
    from algonauts2025_b200.model import FmriEncoderConfig
    from algonauts2025_b200.pl_module import BrainModule
    from algonauts2025_b200.segment import DevicePrefetcher, synthetic_batch
    from algonauts2025_b200.trainer import MiniTrainer, default_optimizer

    B, K, W = args.batch, steps, warmup
    torch.manual_seed(seed)  # DP: identical seeds on every rank -> identical weights and dropout masks (main.py:492-495)
    cfg = FmriEncoderConfig(n_subjects=4, modality_dropout=0.3, feature_aggregation="cat", layer_aggregation="cat", contrastive_enabled=contrastive)
    model = cfg.build(feature_dims=FEATURE_DIMS, n_outputs=1000, n_output_timesteps=100)
    model.independent_replica = dp == "none"  # ensemble members train without gradient exchange on purpose
    module = BrainModule(model=model, loss=torch.nn.MSELoss(), optim_config=None, metrics={}, max_epochs=15)
    # the reference recipe (Adam lr 1e-4 + OneCycleLR per step); the stock torch.optim.Adam instance is adopted by the
    # fused Adam + bf16-shadow kernel exactly as BrainModule.configure_optimizers does (--stock-adam keeps torch's)
    opt, sched = default_optimizer(model.parameters(), total_steps=4 * (K + W) + 16, model=None if args.stock_adam else model)
    # N > 1: our NVLink step tail (copy-engine gradient pushes during the backward -> one fused reduce / rank-sharded Adam /
    # bf16-shadow-multicast kernel per owned range, parallel.ShardedStep); --dp allreduce = NCCL all-reduce + replicated Adam
    sync = None
    if world > 1 and dp != "none":
        if dp == "allreduce" or args.stock_adam:
            sync = parallel.GradAllReduce(model, gemm_sms_during_comm=args.comm_gemm_sms)
        else:
            sync = parallel.data_parallel(model, opt, max_blocks=args.dp_blocks, mode=dp)
    sharded = isinstance(sync, parallel.ShardedStep)
    # whole-step CUDA graphs; NCCL all-reduces inside a step are only captured on request, our own step tail always is
    use_graphs = not args.eager and not args.stock_adam and (sync is None or args.graph_comm or sharded)
    trainer = MiniTrainer(module, opt, sched, grad_sync=sync, use_graphs=use_graphs, graph_collectives=args.graph_comm or sharded,
                          overlap_optimizer=args.overlap and not args.stock_adam,
                          fuse_optimizer=None if args.fused_adam < 0 else bool(args.fused_adam))

    # two distinct host batches (pinned) alternate through the two persistent device slots of a DevicePrefetcher;
    # 210 MB of features per step > L2 (126 MB), so inputs never sit in L2
    data_seed = 1234 + (17 * rank if dp != "none" else 0)  # ensemble members see the same data stream (run_ensemble.py)
    host = [synthetic_batch(batch_size=B, seed=data_seed + i, pin=True) for i in range(2)]
    prefetch = DevicePrefetcher(())
    dev = prefetch.resident(host)
    h2d = sum(v.numel() * v.element_size() for v in host[0].data.values())

    for i in range(W):
        trainer.eager_step(dev[i % 2])
    _barrier(world)
    n_graphs = 0
    if use_graphs:
        # steady state of a long run: every modality-dropout variant of the step has been captured (captures execute
        # nothing; weights, optimizer state and RNG streams are untouched)
        n_graphs = trainer._graphed.warm(dev)
        for i in range(2):
            trainer.train_step(dev[i % 2])
        _barrier(world)

    # ---- timed region 1: device-resident inputs
    launches0 = algonauts2025_b200.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _barrier(world)
    e0.record()
    t_cpu = time.perf_counter()
    for i in range(K):
        trainer.train_step(dev[i % 2])
    cpu_enqueue_ms = 1e3 * (time.perf_counter() - t_cpu) / K  # host time to enqueue one step (no sync inside the loop)
    e1.record()
    _barrier(world)
    t_end = time.perf_counter()
    launches = algonauts2025_b200.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    clk = clocks.stop(t_cpu, t_end) if clocks is not None else None

    # ---- timed region 2: end to end through the public API from pinned host memory
    ms_e2e, last = float("nan"), float("nan")
    if e2e:
        _barrier(world)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        # public-API loop a user writes: pinned host batches -> DevicePrefetcher (H2D of batch i+1 overlaps step i) -> train_step
        # every step's loss is read back to pinned host memory inside the timed region (asynchronously, like a logger that
        # consumes it later; a blocking .item() per step would only measure Python launch latency after each sync)
        loss_host = torch.empty(K, dtype=torch.float32).pin_memory()
        for i, batch in enumerate(prefetch.feed(host[j % 2] for j in range(K))):
            loss_host[i].copy_(trainer.train_step(batch), non_blocking=True)
        f1.record()
        _barrier(world)
        ms_e2e = f0.elapsed_time(f1)
        last = float(loss_host[-1])
        assert all(math.isfinite(float(x)) for x in loss_host), "non-finite loss in the end-to-end loop"

    # ---- roofline pass: K more steps with a CUDA-event pair around every tcgen05 GEMM launch.  With graphs the pairs
    # are captured as event-record nodes of an instrumented copy of the (no-modality-dropped) step graph and read back
    # after each replay, so the GEMMs are timed inside a GPU-bound step exactly like in the timed region.
    gemm_ms, gemm_flops, ms_instr = 0.0, 0.0, float("nan")
    if roofline:
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if use_graphs:
            entries = []
            for j in range(2):
                ops.GEMM_LOG = []
                entry = trainer._graphed._capture(dev[j], [[] for _ in range(2 if contrastive else 1)])
                entries.append((entry, ops.GEMM_LOG))
                ops.GEMM_LOG = None
            _barrier(world)
            g0.record()
            for i in range(K):
                entry, log = entries[i % 2]
                trainer._graphed._replay(entry)
                torch.cuda.synchronize()
                gemm_ms += sum(a.elapsed_time(b) for a, b, _ in log)
                gemm_flops += sum(f for _, _, f in log)
            g1.record()
        else:
            ops.GEMM_LOG = []
            g0.record()
            for i in range(K):
                trainer.eager_step(dev[i % 2])
            g1.record()
            _barrier(world)
            gemm_ms = sum(a.elapsed_time(b) for a, b, _ in ops.GEMM_LOG)
            gemm_flops = sum(f for _, _, f in ops.GEMM_LOG)
            ops.GEMM_LOG = None
        _barrier(world)
        ms_instr = g0.elapsed_time(g1)
    if sharded:
        sync.check()  # no cross-rank wait of the step tail timed out

    ms, ms_e2e, gemm_ms, ms_instr = _max_over_ranks([ms, ms_e2e if e2e else 0.0, gemm_ms, ms_instr if roofline else 0.0], world)
    peaks = measured_peaks()
    n_models = world
    out = {"value": n_models * B * K / (ms / 1e3), "ms_per_step": ms / K, "steps": K, "warmup": W, "contrastive": contrastive,
           "gpu_launches": launches, "host_enqueue_ms_per_step": cpu_enqueue_ms, "last_loss": last, "h2d": h2d,
           "cuda_graphs": {"enabled": use_graphs, "variants_captured": n_graphs, "replays": trainer._graphed.replays if use_graphs else 0},
           "gradient_path": ("none (1 GPU)" if world == 1 else "none (independent ensemble members)" if dp == "none" else
                             sync.describe() if sharded else "NCCL all-reduce (fp32, 9 buckets) + replicated fused Adam"),
           "optimizer_in_backward": bool(trainer.fuse_optimizer), "clocks": clk}
    if e2e:
        out["e2e"] = {"value": n_models * B * K / (ms_e2e / 1e3), "unit": "windows/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4}
    step_tflops = algorithmic_flops_per_window(contrastive) * B * K / (ms / 1e3) / 1e12
    out["whole_step_tflops"], out["whole_step_frac"] = step_tflops, step_tflops / peaks["tflops"]
    if roofline:
        achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
        ref = ncu_reference()
        out["roofline"] = {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                           # context from ncu captures, with provenance (profiles/ncu_reference.json); NOT measured by this run
                           "traffic": (ref.get("train_step_gemm_dram_bytes", {}).get("value") if not contrastive else None),
                           "traffic_unit": "dram bytes per train step, all GEMM launches", "traffic_source": ref.get("train_step_gemm_dram_bytes", {}).get("source"),
                           "tensor_pipe_active_pct_ncu": ref.get("tensor_pipe_active_pct", {}).get("value"),
                           "tensor_pipe_source": ref.get("tensor_pipe_active_pct", {}).get("source"),
                           "kernel": "gemm2_bf16_kernel / gemm_bf16_kernel (all tcgen05 GEMM launches of the timed steps)",
                           "peak_source": peaks["src"] + " sustained bf16", "gemm_share_of_step": gemm_ms / (ms_instr if ms_instr else ms),
                           "measured_in": ("K replays of an instrumented step graph (event-record nodes around every GEMM launch)" if use_graphs
                                           else "the K eager steps, CUDA-event pair per GEMM launch"), "instrumented_ms_per_step": ms_instr / K,
                           "whole_step_tflops": step_tflops, "whole_step_frac": step_tflops / peaks["tflops"]}
    del trainer, opt, sched, module, dev, prefetch
    if sharded:
        model._engine.comm = None
    if not keep_model:
        del model
        model = None
    torch.cuda.empty_cache()
    return out, model


def ensemble_eval_leg(model, rank, world, B):
    """BASELINE config 4, evaluation half (average_submissions.py:107-125): every member predicts the same windows, its
    per-parcel validation r weights its predictions, one all-reduce forms the ensemble prediction, the Pearson kernel
    scores it."""
    from algonauts2025_b200 import ops, parallel
    from algonauts2025_b200.segment import synthetic_batch

    model.eval()
    shared = [synthetic_batch(batch_size=B, seed=999 + i) for i in range(4)]
    shared = [type(b)(data={k: v.cuda() for k, v in b.data.items()}, segments=b.segments) for b in shared]
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def evaluate():
        preds = torch.empty(4 * B, 1000, 100, device="cuda")
        for i, b in enumerate(shared):
            model(b, out=preds[i * B:(i + 1) * B])                              # this member's predictions, written in place
        trues = torch.cat([b.data["fmri"] for b in shared])
        r_member = ops.pearson_r(preds, trues, layout="bdt")[0]                # (O,) per-parcel r of this member
        ens = parallel.ensemble_average(preds, r_member, temperature=0.3)      # weighted all-reduce (reference weighting)
        _, mean_ens = ops.pearson_r(ens.contiguous(), trues, layout="bdt", want_mean=True)
        return r_member, mean_ens

    with torch.no_grad():
        evaluate()  # warm-up pass: NCCL sets up its channels for this message size on first use (0.2 s once)
        _barrier(world)
        f0.record()
        r_member, mean_ens = evaluate()
        f1.record()
    _barrier(world)
    (ms_eval,) = _max_over_ranks([f0.elapsed_time(f1)], world)
    model.train()
    return {"windows": 4 * B, "members": world, "ms": ms_eval, "windows_per_s": 4 * B / (ms_eval / 1e3), "mean_r_member0": float(r_member.mean()),
            "mean_r_ensemble": float(mean_ens[0]), "collective": "weighted all-reduce (NCCL) of the members' predictions inside the timed region",
            "weights": "softmax(r / 0.3) over the voxel axis per member (average_submissions.py:108-109)"}


def eval_sweep_leg(model, rank, world, n_windows: int = EVAL_WINDOWS, batch: int = 64):
    """BASELINE config 5 end to end: every rank predicts ITS shard of the held-out windows from pinned host batches
    (H2D inside the timed region), one all-to-all re-lays (window shard x all parcels) into (all windows x parcel shard),
    every rank reduces 1000/G parcels with the Pearson kernel, r is gathered and read back — the public call
    ``metrics.compute_multidim_pearson(model, loader, distributed="parcels")`` (main.py:459-477)."""
    from algonauts2025_b200 import metrics
    from algonauts2025_b200.segment import DevicePrefetcher, synthetic_batch

    per_rank = n_windows // world
    n_batches = max(1, per_rank // batch)
    per_rank = n_batches * batch
    host = [synthetic_batch(batch_size=batch, seed=4321 + 31 * rank + i, pin=True) for i in range(2)]
    prefetch = DevicePrefetcher(())
    model.eval()
    # warm-up = one full-size pass: NCCL connects the channels of a peer pair lazily, and a 256 MB all-to-all uses more of
    # them than a small one (measured: first large exchange 0.4-4 s, steady state 0.8 ms; tools/a2a_probe.py)
    metrics.compute_multidim_pearson(model, prefetch.feed(host[j % 2] for j in range(n_batches)), distributed="parcels", n_windows=per_rank)
    best = None
    for _ in range(2):  # two timed passes, the faster one is reported (the set is small: 2560 windows = 0.2 s at 8 GPUs)
        _barrier(world)
        timings = {}
        t0 = time.perf_counter()
        r = metrics.compute_multidim_pearson(model, prefetch.feed(host[j % 2] for j in range(n_batches)), distributed="parcels",
                                             n_windows=per_rank, timings=timings)
        wall = time.perf_counter() - t0
        torch.cuda.synchronize()
        (wall_max,) = _max_over_ranks([wall], world)
        if best is None or wall_max < best[0]:
            best = (wall_max, wall, timings, r)
    _, wall, timings, r = best
    assert r.shape == (EVAL_PARCELS,) and np_isfinite(r)
    stage = {k: v[0].elapsed_time(v[-1]) for k, v in timings.items() if len(v) >= 2}
    vals = _max_over_ranks([wall] + [stage.get(k, 0.0) for k in ("predict", "exchange", "pearson", "gather")], world)
    wall, stage_ms = vals[0], dict(zip(("predict", "exchange", "pearson", "gather"), vals[1:]))
    model.train()
    total = per_rank * world
    h2d = sum(v.numel() * v.element_size() for v in host[0].data.values())
    return {"metric": "eval sweep (predict -> parcel all-to-all -> sharded Pearson -> gather) windows/s", "value": total / wall, "unit": "windows/s",
            "parcel_trs_per_s": total * EVAL_PARCELS * EVAL_TRS / wall, "windows": total, "windows_per_gpu": per_rank, "batch_per_gpu": batch,
            "wall_s": wall, "stage_ms_max_over_ranks": stage_ms, "mean_r": float(r.mean()),
            "e2e": {"value": total / wall, "unit": "windows/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * EVAL_PARCELS,
                    "api": 'metrics.compute_multidim_pearson(model, DevicePrefetcher(pinned host batches), distributed="parcels") -> r[1000] on host'},
            "exchange_bytes_per_gpu": 2 * 4 * per_rank * EVAL_TRS * (EVAL_PARCELS - EVAL_PARCELS // world) if world > 1 else 0,
            "collectives": ("NCCL all-to-all of predictions + targets, all-gather of r — inside the timed region" if world > 1 else "none (1 GPU)")}


def np_isfinite(a):
    import numpy as np

    return bool(np.isfinite(a).all())


def run_ours(args):
    import torch.distributed as dist

    import algonauts2025_b200

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    algonauts2025_b200.load()
    B, K, W = args.batch, args.steps, args.warmup
    contrastive = bool(args.contrastive)
    extras = not args.headline_only

    # The Pearson-eval leg is an independent metric of a bandwidth kernel "timed alone": it runs FIRST, before the train legs
    # put the board under its power cap (tools/pearson_variance_probe.py: 6.1-6.3 TB/s on a cool chip, 5.5 TB/s right after
    # 20 s of GEMM load with the SM clock at 1342 MHz — the HBM peak in MEASURED_PEAKS.json is a cool-chip figure too).
    pe = pearson_eval_leg(rank, world, max(K, 5), W, measured_peaks(), world == 1 and not args.no_cpu_baseline) if not args.no_pearson else None
    torch.cuda.empty_cache()

    clocks = ClockSampler(local)
    clocks.start()  # nvidia-smi needs a few hundred ms to deliver its first sample: start ahead of the timed region
    main_leg, model = train_leg(args, rank, world, local, contrastive=contrastive, dp=args.dp, seed=33, steps=K, warmup=W, roofline=True, e2e=True,
                                clocks=clocks, keep_model=True)
    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_steps(2, 1, REF_WINDOWS_PER_STEP, contrastive)  # bounded sample: ~20 s of host work (1.7 TFLOP per window-step)
            cpu = {"value": r["windows_per_s"], "unit": "windows/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        line = {"metric": "train windows/s", "value": main_leg["value"], "unit": "windows/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": main_leg["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": WORKLOAD, "global_batch": world * B, "contrastive": contrastive, "parallelism": f"dp{world}",
                           "gradient_path": main_leg["gradient_path"],
                           "optimizer": ("torch Adam(fused)" if args.stock_adam else "Adam (fused Adam+bf16-shadow kernel" + (", rank-sharded" if world > 1 and args.dp not in ("allreduce", "none") else "") + ")") + " + OneCycleLR, fp32 master weights",
                           "optimizer_overlap": bool(args.overlap), "optimizer_in_backward": main_leg["optimizer_in_backward"], "l2": "inputs larger than L2 (210 MB features/step, 2 alternating batches)",
                           "last_loss": main_leg["last_loss"], "host_enqueue_ms_per_step": main_leg["host_enqueue_ms_per_step"],
                           "cuda_graphs": main_leg["cuda_graphs"]},
                "e2e": main_leg["e2e"], "gpu_launches": main_leg["gpu_launches"], "roofline": main_leg["roofline"], "clocks": main_leg["clocks"]}
        if cpu is not None:
            line["cpu_baseline"] = cpu

    # ---- BASELINE config 5 (eval sweep) on the trained replica: batched predict + parcel-sharded Pearson
    ev = sweep = None
    if not args.no_pearson:
        ev = eval_predict_leg(model, rank, world, max(K // 2, 4))
        if extras:
            sweep = eval_sweep_leg(model, rank, world)
    del model
    torch.cuda.empty_cache()

    # ---- the reference's default recipe has the contrastive branch ON (defaults.py:102): second encoder pass + InfoNCE
    con = ens = None
    if extras and not contrastive:
        con, _ = train_leg(args, rank, world, local, contrastive=True, dp=args.dp, seed=33, steps=max(K // 2, 5), warmup=max(W, 3), roofline=True, e2e=False)
    # ---- BASELINE config 4 (run_ensemble): one independent member per GPU (own seed, no training collectives) + ensemble eval
    if extras and world > 1:
        ens, member = train_leg(args, rank, world, local, contrastive=False, dp="none", seed=1000 + rank, steps=max(K // 2, 5), warmup=max(W, 3),
                                roofline=False, e2e=False, keep_model=True)
        ens["ensemble_eval"] = ensemble_eval_leg(member, rank, world, B)
        del member
        torch.cuda.empty_cache()

    if rank == 0:
        if ev is not None:
            line["eval_predict"] = ev
        if sweep is not None:
            line["eval_sweep"] = sweep
        if pe is not None:
            line["pearson_eval"] = pe
        if con is not None:
            line["contrastive_on"] = {"metric": "train windows/s, contrastive branch on (reference default, defaults.py:102)", "value": con["value"], "unit": "windows/s",
                                      "ms_per_step": con["ms_per_step"], "steps": con["steps"], "gpu_launches": con["gpu_launches"], "cuda_graphs": con["cuda_graphs"],
                                      "gradient_path": con["gradient_path"], "roofline": con["roofline"]}
        if ens is not None:
            line["ensemble"] = {"metric": "ensemble train windows/s (BASELINE config 4: one member per GPU, no training collectives)", "value": ens["value"],
                                "unit": "windows/s", "members": world, "ms_per_step": ens["ms_per_step"], "steps": ens["steps"], "scaling": "weak",
                                "whole_step_frac": ens["whole_step_frac"], "gpu_launches": ens["gpu_launches"], "ensemble_eval": ens["ensemble_eval"]}
        elif extras:
            line["ensemble"] = {"skipped": "1 GPU: an ensemble of one member is the headline train step; run with --gpus >= 2"}
        print(json.dumps(line), flush=True)
    if world > 1:
        _barrier(world)
        if args.graph_comm and args.dp == "allreduce":
            # CUDA graphs that captured NCCL collectives keep the communicator busy: destroy_process_group() blocks on
            # them (seen on 2 x B200).  The result is printed and every rank has passed the barrier — leave directly.
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)
        # our own step tail puts no NCCL work into the graphs: a normal teardown (bounded, in case a communicator is stuck)
        done = threading.Event()

        def _teardown():
            dist.destroy_process_group()
            done.set()

        threading.Thread(target=_teardown, daemon=True).start()
        if not done.wait(30.0):
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)


def run_ensemble(args):
    """``--mode ensemble``: only the BASELINE config 4 legs (the default run reports them under the ``ensemble`` key)."""
    import torch.distributed as dist

    import algonauts2025_b200

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    algonauts2025_b200.load()
    ens, member = train_leg(args, rank, world, local, contrastive=False, dp="none", seed=1000 + rank, steps=args.steps, warmup=args.warmup,
                            roofline=False, e2e=False, keep_model=True)
    ev = ensemble_eval_leg(member, rank, world, args.batch)
    if rank == 0:
        print(json.dumps({"metric": "ensemble train windows/s (one member per GPU)", "value": ens["value"], "unit": "windows/s",
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ens["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                          "config": {"workload": WORKLOAD, "members": world, "parallelism": f"ensemble x{world} (no training collectives)"},
                          "gpu_launches": ens["gpu_launches"], "ensemble_eval": ev}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="train", choices=["train", "ensemble"], help="ensemble: BASELINE config 4, one member per GPU + ensemble-averaged Pearson eval")
    ap.add_argument("--contrastive", type=int, default=0)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pearson", action="store_true", help="skip the evaluation legs (Pearson kernel, eval predict, eval sweep)")
    ap.add_argument("--headline-only", action="store_true", help="skip the extra legs (contrastive-on, ensemble, eval sweep)")
    ap.add_argument("--eager", action="store_true", help="launch every step from Python instead of replaying whole-step CUDA graphs")
    ap.add_argument("--graph-comm", action="store_true", help="N > 1: capture the steps including their NCCL all-reduces (default: eager steps)")
    ap.add_argument("--dp", default="staged", choices=["staged", "nvls", "p2p", "allreduce", "none"],
                    help="N > 1 gradient path: copy-engine pushes + our fused reduce/Adam/multicast kernel (default), the same kernel reducing "
                         "through NVLS multimem.ld_reduce or peer pointers, or NCCL all-reduce + replicated Adam")
    ap.add_argument("--dp-blocks", type=int, default=0, help="N > 1: CTAs of the sharded step-tail kernel (0 = 6 per SM)")
    ap.add_argument("--comm-gemm-sms", type=int, default=0, help="N > 1: SMs the backward GEMMs use while gradient all-reduces are in flight (0 = all)")
    ap.add_argument("--watchdog", type=int, default=1500, help="dump all thread stacks and exit if the run takes longer than this many seconds (0 = off)")
    ap.add_argument("--overlap", action="store_true", help="run each layer's Adam step behind the backward (parallel.StepOverlap) instead of after it")
    ap.add_argument("--fused-adam", type=int, default=-1, help="1/0: Adam update of the GEMM weights inside the wgrad GEMM epilogues (optimizer-in-backward); "
                    "-1 = on whenever nothing sits between backward and step (1 GPU, ensemble members)")
    ap.add_argument("--stock-adam", action="store_true", help="keep torch's multi-tensor fused Adam instead of the TribeAdam kernel")
    args = ap.parse_args()
    if args.watchdog > 0:
        import faulthandler

        faulthandler.dump_traceback_later(args.watchdog, exit=True)
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
        run_ensemble(args) if args.mode == "ensemble" else run_ours(args)


if __name__ == "__main__":
    main()
