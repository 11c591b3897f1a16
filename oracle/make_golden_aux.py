"""TEST INFRASTRUCTURE — generates ``tests/golden/aux_ops.npz`` by running the REFERENCE's own code for the steps either
side of the hot path.  Build container only (needs /root/reference):  ``python -m oracle.make_golden_aux``

Executed unmodified from /root/reference (absent third-party packages satisfied by ``oracle/shims``):
``modeling_utils.losses.losses.PearsonLoss`` (+ autograd gradient), ``modeling_utils.metrics.metrics.Rank / TopkAcc``,
``data_utils.base.TimedArray`` in the extractors' two-stage flow (features/audio.py:100-111, 236-252),
``data_utils.segments._prepare_strided_windows``, ``algonauts2025.grids.average_submissions.average_submissions`` on
synthetic submission folders and the ``algonauts2025.callbacks.Benchmark`` callback on mock segments.  torch.nn.SmoothL1Loss / HuberLoss / L1Loss are the stock torch modules the reference's
``TorchLossConfig`` instantiates (losses/base.py:43-59)."""
from __future__ import annotations

import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle", "shims"), REF, f"{REF}/data_utils", f"{REF}/modeling_utils"]

from algonauts2025.grids import average_submissions as A  # noqa: E402
from data_utils.base import TimedArray  # noqa: E402
from data_utils.segments import _prepare_strided_windows  # noqa: E402
from modeling_utils.losses.losses import PearsonLoss  # noqa: E402
from modeling_utils.metrics.metrics import Rank, TopkAcc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def ref_window(arr, arr_start, f, win_start, win_duration):
    tdata = TimedArray(data=arr, start=arr_start, frequency=f)
    sub = tdata.overlap(start=win_start, duration=win_duration)
    if sub is None:
        sub = tdata.overlap(start=tdata.start, duration=0)
    out = TimedArray(aggregation="sum", start=win_start, frequency=f, duration=win_duration)
    out += sub
    return out.data


def main():
    out = {}
    g = torch.Generator().manual_seed(17)
    # ---- losses
    p = torch.randn(96, 13, generator=g)
    t = 0.4 * p + torch.randn(96, 13, generator=g)
    out["loss_pred"], out["loss_true"] = p.numpy(), t.numpy()
    for red in ("mean", "sum"):
        x = p.clone().requires_grad_(True)
        val = PearsonLoss(reduction=red)(x, t)
        val.backward()
        out[f"pearson_{red}"], out[f"pearson_{red}_grad"] = val.detach().numpy(), x.grad.numpy()
    for name, mod in (("smooth_l1", torch.nn.SmoothL1Loss()), ("smooth_l1_b05", torch.nn.SmoothL1Loss(beta=0.5)), ("huber", torch.nn.HuberLoss()),
                      ("huber_d2", torch.nn.HuberLoss(delta=2.0)), ("l1", torch.nn.L1Loss())):
        x = (2.0 * p).clone().requires_grad_(True)
        val = mod(x, t)
        val.backward()
        out[f"{name}"], out[f"{name}_grad"] = val.detach().numpy(), x.grad.numpy()
    # ---- retrieval ranks (TopkAcc as val/retrieval_top1: pl_module.py:100-101 feeds time-averaged (B, O) tensors)
    x = torch.randn(16, 40, generator=g)
    y = 0.6 * x + torch.randn(16, 40, generator=g)
    y[5] = y[2]                      # duplicate candidate -> tie handling
    x[9, 3] = float("nan")           # NaN query -> rank falls back to n // 2
    out["rank_x"], out["rank_y"] = x.numpy(), y.numpy()
    m1, m5, mr = TopkAcc(topk=1), TopkAcc(topk=5), Rank(reduction="mean")
    for metric in (m1, m5, mr):
        metric.update(x, y)
        metric.update(x[:7], y[:7])
    out["rank_ranks"] = m1.ranks.numpy()
    out["rank_top1"], out["rank_top5"] = m1.compute().numpy(), m5.compute().numpy()
    out["rank_scores"] = Rank._compute_sim(x, y).numpy()
    # ---- windows: (n_samples, arr_start, frequency, win_start, win_duration)
    cases = []
    rng = np.random.default_rng(3)
    for n, a0, f in ((700, 0.0, 2.0), (1233, 1.37, 2.0), (290, 10.0, 2.0), (469, 0.0, 1 / 1.49), (97, 4.47, 1 / 1.49), (5, 300.0, 2.0)):
        dur_total = n / f
        starts = list(np.arange(a0 - 4.47, a0 + dur_total - 4.47 + 1e-8, 149.0)) + [a0 - 200.0, a0 + dur_total - 0.2, a0 + 0.26, a0 - 148.9]
        for ws in starts:
            cases.append((n, a0, f, float(ws), 149.0))
    cases.append((400, 0.0, 2.0, 12.25, 30.0))
    cases.append((400, 0.0, 2.0, 12.75, 30.5))
    wins = []
    for i, (n, a0, f, ws, wd) in enumerate(cases):
        arr = (rng.standard_normal((2, n)) + np.arange(n)[None, :]).astype(np.float32)
        out[f"win_arr_{i}"] = arr
        out[f"win_out_{i}"] = ref_window(arr, a0, f, ws, wd)
        wins.append((n, a0, f, ws, wd))
    out["win_cases"] = np.array(wins, dtype=np.float64)
    for i, (a, b) in enumerate(((-4.47, 700 - 4.47), (-4.47 + 3.3, 1490.0 - 4.47 + 3.3), (10.0, 10.0), (0.0, 148.99999999))):
        s, d = _prepare_strided_windows(a, b, 149.0, 149.0, drop_incomplete=False)
        out[f"stride_in_{i}"], out[f"stride_starts_{i}"], out[f"stride_durs_{i}"] = np.array([a, b]), s, d
    # ---- ensemble averaging through the reference function on synthetic submission folders
    M, O = 4, 24
    members = [{"sub-01": {"c1": rng.standard_normal((9, O)).astype(np.float32), "c2": rng.standard_normal((5, O)).astype(np.float32)},
                "sub-02": {"c1": rng.standard_normal((3, O)).astype(np.float32)}} for _ in range(M)]
    pearsons = rng.uniform(0.0, 0.4, (M, O)).astype(np.float32)
    scores = (0.2 + 0.03 * rng.standard_normal(M)).astype(np.float64)
    out["ens_pearsons"], out["ens_scores"] = pearsons, scores
    for m in range(M):
        for sub, chunks in members[m].items():
            for c, v in chunks.items():
                out[f"ens_in_{m}_{sub}_{c}"] = v
    import pandas as pd

    for tag, kw in (("voxel", dict(weigh_by_score=True, per_voxel_weights=True, temperature=0.3)),
                    ("scalar", dict(weigh_by_score=True, per_voxel_weights=False, temperature=0.05)), ("mean", dict(weigh_by_score=False))):
        with tempfile.TemporaryDirectory() as d:
            d = Path(d)
            for m in range(M):
                f = d / f"run{m}"
                f.mkdir()
                np.savez(str(f / "submission"), submission=members[m])
                os.rename(f / "submission.npz", f / "submission.zip")
                pd.DataFrame({"val/pearson": [scores[m]]}).to_csv(f / "metrics.csv")
                np.save(f / "pearson.npy", pearsons[m])
            A.average_submissions(d, **kw)
            res = np.load(d / "submission.npy", allow_pickle=True).item()
            for sub, chunks in res.items():
                for c, v in chunks.items():
                    out[f"ens_{tag}_{sub}_{c}"] = np.asarray(v)
    # ---- submission assembly through the reference Benchmark callback (callbacks.py:47-103) on mock segments
    import types

    from algonauts2025.callbacks import Benchmark

    class _Seg:
        def __init__(self, subject, chunk):
            self.events = pd.DataFrame({"subject": [subject] * 2, "chunk": [chunk] * 2})

    # NB callbacks.py:59 sets ``overlap_trs = 0.0`` (a float), so ``pred[overlap_trs:]`` (:73) raises TypeError as soon as a
    # (subject, chunk) pair receives a SECOND window; the reference callback therefore only runs with one window per
    # chunk, which is what these vectors cover.  Multi-window chunks are checked against the oracle (overlap 0).
    layout = [[("Algonauts2025/sub-01", "friends:e01a"), ("Algonauts2025/sub-02", "friends:e01a"), ("Algonauts2025/sub-01", "friends:e01b")],
              [("Algonauts2025/sub-02", "friends:e01b"), ("Algonauts2025/sub-01", "friends:e02a"), ("Algonauts2025/sub-02", "friends:e02a")]]
    samples = {"sub-01": {"s07e01a": 4, "s07e01b": 5, "s07e02a": 3}, "sub-02": {"s07e01a": 5, "s07e01b": 2, "s07e02a": 4}}
    with tempfile.TemporaryDirectory() as d:
        d = Path(d)
        for subj, per in samples.items():
            f = d / f"algonauts_2025.competitors/fmri/{subj}/target_sample_number"
            f.mkdir(parents=True)
            np.save(f / f"{subj}_friends-s7_fmri_samples.npy", per)
        cb = Benchmark(d)
        trainer = types.SimpleNamespace(logger=types.SimpleNamespace(save_dir=str(d)))
        cb.on_test_epoch_start(trainer, None)
        for b, segs in enumerate(layout):
            y = torch.randn(3, 6, 5, generator=g)
            out[f"sub_pred_{b}"] = y.numpy()
            batch = types.SimpleNamespace(segments=[_Seg(s, c) for s, c in segs])
            cb.on_test_batch_end(trainer, None, (y, None), batch, b)
        cb.on_test_epoch_end(trainer, None)
        for subj, per in cb.submission_dict.items():
            for chunk, arr in per.items():
                out[f"sub_out_{subj}_{chunk}"] = np.asarray(arr)
    out["sub_layout"] = np.array([[f"{s}|{c}" for s, c in segs] for segs in layout])
    np.savez_compressed(os.path.join(OUT, "aux_ops.npz"), **out)
    print("aux_ops.npz written:", len(out), "arrays")


if __name__ == "__main__":
    main()
