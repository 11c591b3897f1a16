"""CPU: the §8f oracle restatements (oracle/aux_oracle.py) and the product's host-side window index arithmetic
(algonauts2025_b200/windows.py) against vectors produced by RUNNING the reference (tests/golden/aux_ops.npz,
oracle/make_golden_aux.py).  Index work is bit-exact; float work within 1e-6."""
import os

import numpy as np
import pytest

import algonauts2025_b200  # noqa: F401
from algonauts2025_b200 import windows as W
from oracle import aux_oracle as A


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "aux_ops.npz"))


def test_pearson_loss_oracle_matches_reference_value_and_grad(g):
    for red in ("mean", "sum"):
        val, grad = A.pearson_loss_and_grad(g["loss_pred"], g["loss_true"], red)
        np.testing.assert_allclose(val, g[f"pearson_{red}"], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(grad, g[f"pearson_{red}_grad"], rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("name,kind,param", [("smooth_l1", "smooth_l1", 1.0), ("smooth_l1_b05", "smooth_l1", 0.5), ("huber", "huber", 1.0),
                                             ("huber_d2", "huber", 2.0), ("l1", "l1", 0.0)])
def test_point_losses_oracle_matches_torch_modules(g, name, kind, param):
    val, grad = A.point_loss_and_grad(2.0 * g["loss_pred"], g["loss_true"], kind, param)
    np.testing.assert_allclose(val, g[name], rtol=1e-6)
    np.testing.assert_allclose(grad, g[f"{name}_grad"], rtol=1e-5, atol=1e-9)


def test_retrieval_ranks_oracle_matches_reference(g):
    x, y = g["rank_x"], g["rank_y"]
    np.testing.assert_allclose(A.retrieval_scores(x, y), g["rank_scores"], rtol=1e-5, atol=1e-6, equal_nan=True)
    ranks = np.concatenate([A.retrieval_ranks(x, y), A.retrieval_ranks(x[:7], y[:7])])
    np.testing.assert_array_equal(ranks, g["rank_ranks"])
    assert ranks[9] == 16 // 2                      # NaN query
    assert ranks[2] == ranks[5] or ranks[2] % 1 == 0.5  # duplicated candidate: tie averaged
    assert A.topk_acc(ranks, 1) == pytest.approx(float(g["rank_top1"]))
    assert A.topk_acc(ranks, 5) == pytest.approx(float(g["rank_top5"]))


def test_window_assembly_oracle_and_product_indices_are_bit_exact(g):
    cases = g["win_cases"]
    seen_empty = seen_partial = seen_full = 0
    for i, (n, a0, f, ws, wd) in enumerate(cases):
        arr, ref = g[f"win_arr_{i}"], g[f"win_out_{i}"]
        np.testing.assert_array_equal(A.assemble_window(arr, a0, f, ws, wd), ref)
        d0, s0, ln, t_win = W.window_triple(ws, wd, f, a0, int(n))       # the product's host arithmetic
        mine = np.zeros_like(ref)
        assert t_win == ref.shape[-1]
        mine[..., d0: d0 + ln] = arr[..., s0: s0 + ln]
        np.testing.assert_array_equal(mine, ref)
        seen_empty += ln == 0
        seen_partial += 0 < ln < t_win
        seen_full += ln == t_win
    assert seen_empty and seen_partial and seen_full


def test_strided_windows_match_reference(g):
    for i in range(4):
        a, b = g[f"stride_in_{i}"]
        starts, durs = W.strided_windows(a, b, 149.0, 149.0, drop_incomplete=False)
        np.testing.assert_array_equal(starts, g[f"stride_starts_{i}"])
        np.testing.assert_array_equal(durs, g[f"stride_durs_{i}"])
        np.testing.assert_array_equal(A.strided_window_starts(a, b), g[f"stride_starts_{i}"])
    s, _ = W.timeline_windows(0.0, 700.0)
    np.testing.assert_array_equal(s, g["stride_starts_0"])


def test_ensemble_oracle_matches_reference_average_submissions(g):
    M = g["ens_pearsons"].shape[0]
    for tag, kw in (("voxel", dict(weigh_by_score=True, per_voxel_weights=True, temperature=0.3)),
                    ("scalar", dict(weigh_by_score=True, per_voxel_weights=False, temperature=0.05)), ("mean", dict(weigh_by_score=False))):
        for sub, chunk in (("sub-01", "c1"), ("sub-01", "c2"), ("sub-02", "c1")):
            preds = np.stack([g[f"ens_in_{m}_{sub}_{chunk}"] for m in range(M)])
            got = A.average_members(preds, g["ens_pearsons"], g["ens_scores"], **kw)
            np.testing.assert_allclose(got, g[f"ens_{tag}_{sub}_{chunk}"], rtol=1e-5, atol=1e-6)


def test_swa_rule():
    rng = np.random.default_rng(0)
    ps = rng.standard_normal((5, 11))
    avg = np.zeros(11)
    for n, p in enumerate(ps):
        avg = A.swa_update(avg, p, n)
    np.testing.assert_allclose(avg, ps.mean(0), rtol=1e-12)


def _submission_inputs(g):
    layout = [[tuple(x.split("|")) for x in row] for row in g["sub_layout"]]
    batches = [g[f"sub_pred_{b}"] for b in range(len(layout))]
    labels = [[(s.split("/")[1], "s07" + c.split(":")[1]) for s, c in row] for row in layout]
    samples = {}
    for k in g.files:
        if k.startswith("sub_out_"):
            _, _, subj, chunk = k.split("_")
            samples.setdefault(subj, {})[chunk] = g[k].shape[0]
    return layout, batches, labels, samples


def test_submission_assembly_oracle_matches_reference_benchmark_callback(g):
    layout, batches, labels, samples = _submission_inputs(g)
    out = A.assemble_submission(batches, labels, samples)
    for subj, per in out.items():
        for chunk, arr in per.items():
            np.testing.assert_array_equal(arr, g[f"sub_out_{subj}_{chunk}"])
    with pytest.raises(ValueError):
        A.assemble_submission(batches, labels, {s: {c: n + 10 for c, n in per.items()} for s, per in samples.items()})
