"""Host-side logic of the drop-in boundary that needs no GPU: config surface, parameter names / init order vs the
reference (golden), RNG consumption of the dropout draw, error behaviour, parcel sharding arithmetic, and the
world_size-2 gloo paths of parallel.py."""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pydantic
import pytest
import torch

import algonauts2025_b200
from algonauts2025_b200 import parallel
from algonauts2025_b200.model import FmriEncoder, FmriEncoderConfig
from algonauts2025_b200.segment import SegmentData, synthetic_batch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL_DIMS = {"text": (2, 96), "audio": (2, 40), "video": (1, 72)}


def test_config_surface_matches_reference():
    cfg = FmriEncoderConfig(n_subjects=4)
    assert cfg.model_dump() == {"name": "FmriEncoder", "n_subjects": 4, "feature_aggregation": "cat", "layer_aggregation": "cat",
                                "subject_embedding": False, "modality_dropout": 0.0, "contrastive_enabled": False,
                                "contrastive_modalities": ["video"], "contrastive_weight": 0.1, "contrastive_temperature": 0.07}
    # the one extension field stays out of the dump (config hashes = the reference's) and is validated
    assert FmriEncoderConfig(n_subjects=4, xt_semantics="v1.27").xt_semantics == "v1.27"
    assert "xt_semantics" not in FmriEncoderConfig(n_subjects=4, xt_semantics="v1.27").model_dump()
    with pytest.raises(pydantic.ValidationError):
        FmriEncoderConfig(n_subjects=4, xt_semantics="v3")
    with pytest.raises(pydantic.ValidationError):
        FmriEncoderConfig(n_subjects=4, bogus=1)  # extra="forbid" (model.py:21)
    with pytest.raises(pydantic.ValidationError):
        FmriEncoderConfig(feature_aggregation="mean")


def test_segment_data_contract():
    b = synthetic_batch(batch_size=2, t=5, t_out=3, n_outputs=7)
    assert set(b.data) == {"text", "audio", "video", "fmri", "subject_id"} and len(b.segments) == 2
    assert b.data["subject_id"].dtype == torch.int64 and b.data["fmri"].shape == (2, 7, 3)
    with pytest.raises(RuntimeError):
        b["text"]
    with pytest.raises(RuntimeError):
        SegmentData(data={"x": torch.zeros(3, 1)}, segments=[None])
    with pytest.raises(ValueError):
        SegmentData(data={}, segments=[])


def test_small_model_names_and_errors():
    torch.manual_seed(0)
    m = FmriEncoder(SMALL_DIMS, 20, 5, FmriEncoderConfig(n_subjects=2, contrastive_enabled=True), hidden=384, depth=1, heads=6)
    names = [n for n, _ in m.named_parameters()]
    assert names[:9] == ["time_pos_embed", "projectors.text.weight", "projectors.text.bias", "projectors.audio.weight",
                         "projectors.audio.bias", "projectors.video.weight", "projectors.video.bias",
                         "contrastive_heads.video.weight", "contrastive_heads.video.bias"]
    assert "encoder.layers.0.0.0.g" in names and "encoder.layers.1.1.ff.2.bias" in names and "encoder.final_norm.g" in names
    assert "encoder.rotary_pos_emb.inv_freq" in m.state_dict()
    assert m.projectors["text"].weight.shape == (128, 192) and m.contrastive_heads["video"].weight.shape == (384, 72)
    assert repr(m.predictor) == "SubjectLayers(384, 20, 2)"
    with pytest.raises(ValueError):
        FmriEncoder(SMALL_DIMS, 20, 5, FmriEncoderConfig(n_subjects=2), hidden=384, depth=1, heads=5)
    with pytest.raises(ValueError):
        FmriEncoder(SMALL_DIMS, 20, 5, FmriEncoderConfig(n_subjects=2), hidden=192, depth=1, heads=6)
    if not torch.cuda.is_available():
        with pytest.raises(algonauts2025_b200.TribeError):  # no CPU fallback
            m(synthetic_batch(batch_size=1, t=8, t_out=5, n_outputs=20, n_subjects=2, dims=(("text", 2, 96), ("audio", 2, 40), ("video", 1, 72))))


@pytest.mark.slow
def test_full_model_names_and_init_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "full_model.npz"))
    torch.manual_seed(33)
    m = FmriEncoder({"text": (2, 3072), "audio": (2, 1024), "video": (2, 1408)}, 1000, 100,
                    FmriEncoderConfig(n_subjects=4, modality_dropout=0.3), device=None if torch.cuda.is_available() else None)
    params = dict(m.named_parameters())
    assert [str(n) for n in g["param_checksum_names"]] == list(params)
    for name, val in zip(g["param_checksum_names"], g["param_checksum_vals"]):
        got = params[str(name)].detach().double().sum().item()
        assert abs(got - val) <= 1e-6 * max(1.0, abs(val)), name


def test_dropout_draws_match_reference_masks(golden_dir):
    meta = json.load(open(os.path.join(golden_dir, "full_model.json")))
    m = FmriEncoder(SMALL_DIMS, 20, 5, FmriEncoderConfig(n_subjects=2), hidden=384, depth=1, heads=6)
    m.train()
    for key, want in meta["dropout_masks"].items():
        if key.endswith("next_rand"):
            continue
        p, seed = key.split(":")
        m.config.modality_dropout = float(p)
        torch.manual_seed(int(seed)), np.random.seed(int(seed))
        assert sorted(m._draw_dropout()) == sorted(want), key
        assert torch.rand(1).item() == meta["dropout_masks"][key + ":next_rand"]


def test_parcel_bounds_cover_and_balance():
    for n, g in ((1000, 1), (1000, 2), (1000, 8), (1003, 8), (5, 8)):
        b = parallel.parcel_bounds(n, g)
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(g - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1


GLOO_WORKER = textwrap.dedent("""
    import os, sys, numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, {root!r})
    import algonauts2025_b200
    from algonauts2025_b200 import parallel
    from oracle import tribe_oracle as O
    rank, ws = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    g = torch.Generator().manual_seed(0)
    preds, trues = torch.randn(50, 13, generator=g), torch.randn(50, 13, generator=g)
    rows = slice(0, 20) if rank == 0 else slice(20, 50)            # ragged window shards
    shard, (lo, hi) = parallel.exchange_parcel_shards(preds[rows].contiguous())
    assert (lo, hi) == parallel.parcel_bounds(13, ws)[rank]
    assert torch.equal(shard, preds[:, lo:hi])                     # all rows, this rank's parcels, source-rank order
    tshard, _ = parallel.exchange_parcel_shards(trues[rows].contiguous())
    r_shard = torch.from_numpy(O.pearson_columns_f64(shard.numpy(), tshard.numpy())).float()
    r = parallel.gather_parcels(r_shard, 13)
    np.testing.assert_allclose(r.numpy(), O.pearson_columns_f64(preds.numpy(), trues.numpy()), atol=1e-6)
    # the same exchange on (windows, parcels, TRs) prediction tensors (the layout FmriEncoder.forward returns)
    p3 = torch.randn(7, 13, 5, generator=g)
    wins = slice(0, 3) if rank == 0 else slice(3, 7)
    shard3, (lo3, hi3) = parallel.exchange_parcel_shards(p3[wins].contiguous())
    assert torch.equal(shard3, p3[:, lo3:hi3])
    # ensemble: one member per rank
    member_pred = preds * (rank + 1)
    member_r = torch.linspace(0.0, 0.3, 13) * (rank + 1)
    ens = parallel.ensemble_average(member_pred, member_r, 0.3, softmax_over="members")
    stack_p = np.stack([preds.numpy(), 2 * preds.numpy()])
    stack_r = np.stack([member_r.numpy() / (rank + 1) * 1, member_r.numpy() / (rank + 1) * 2])
    want = O.ensemble_average(stack_p, stack_r, 0.3)
    np.testing.assert_allclose(ens.numpy(), want, rtol=1e-5, atol=1e-6)
    # default = the reference's arithmetic (average_submissions.py:108-109: softmax over the voxel axis per member)
    from oracle import aux_oracle as AO
    ens_ref = parallel.ensemble_average(member_pred, member_r, 0.3)
    want_ref = AO.average_members(stack_p, stack_r, weigh_by_score=True, per_voxel_weights=True, temperature=0.3)
    np.testing.assert_allclose(ens_ref.numpy(), want_ref, rtol=1e-5, atol=1e-6)
    dist.destroy_process_group()
    print("OK", rank)
""")


def test_parallel_paths_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER.format(root=ROOT))
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"OK {rank}" in out, out[-2000:]
