"""TEST INFRASTRUCTURE — CPU (PyTorch fp32) restatement of TRIBE's encoding-model hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import
this file; the product package (``algonauts-2025_b200``) never does and has no CPU fallback.

What is restated, and where it lives in the reference (paths relative to /root/reference):

* ``OracleFmriEncoder``            <- ``algonauts2025/model.py:46-174``   (init order, aggregate_features incl. the
  CPU-RNG modality dropout, transformer_forward, forward) and ``:177-241`` (contrastive helpers / InfoNCE)
* ``subject_layers``               <- ``modeling_utils/modeling_utils/models/common.py:45-67``
* ``adaptive_pool_windows``        <- ``torch.nn.AdaptiveAvgPool1d`` as used at ``model.py:60,119-122``
* ``run_step``                     <- ``algonauts2025/pl_module.py:46-107``
* ``pearson_loss``                 <- ``modeling_utils/modeling_utils/losses/losses.py:11-42``
* ``multidim_pearson_scipy``       <- ``algonauts2025/main.py:470-477``
* ``ensemble_average``             <- ``algonauts2025/grids/average_submissions.py:107-125``
* encoder / streaming Pearson      <- third-party, see ``oracle/xt_encoder.py`` and ``oracle/tm_pearson.py``

Pinning: ``oracle/make_golden.py`` runs the *reference's own* ``model.py`` / ``pl_module.py`` / ``common.py`` /
``losses.py`` (imported from /root/reference under the shims in ``oracle/shims``) on seeded inputs and writes
``tests/golden/*``; ``tests/test_oracle.py`` checks this restatement against those files.  The reference ships no
tests or golden vectors of its own, and its encoder arithmetic lives in the absent ``x_transformers``, so
the encoder block itself is **parity unpinned** (declared semantics, see ``oracle/xt_encoder.py``).
"""
from __future__ import annotations

import dataclasses
import math
import typing as tp

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from .tm_pearson import PearsonCorrCoef
from .xt_encoder import Encoder

HIDDEN = 3072  # model.py:61


@dataclasses.dataclass
class SegmentData:
    """Stand-in for ``data_utils/dataloader.py:27-53``: ``.data`` dict, ``.segments`` list, ``.to``."""

    data: dict
    segments: list

    def to(self, device):
        return SegmentData({k: v.to(device) for k, v in self.data.items()}, self.segments)


@dataclasses.dataclass
class OracleConfig:
    """Field-for-field ``FmriEncoderConfig`` (model.py:20-33)."""

    n_subjects: int | None = None
    feature_aggregation: str = "cat"
    layer_aggregation: str = "cat"
    subject_embedding: bool = False
    modality_dropout: float = 0.0
    contrastive_enabled: bool = False
    contrastive_modalities: tp.Sequence[str] = ("video",)
    contrastive_weight: float = 0.1
    contrastive_temperature: float = 0.07


def adaptive_pool_windows(t_in: int, t_out: int) -> list[tuple[int, int]]:
    """Window ``i`` of AdaptiveAvgPool1d: ``[floor(i*T/T'), ceil((i+1)*T/T'))``."""
    return [((i * t_in) // t_out, -((-(i + 1) * t_in) // t_out)) for i in range(t_out)]


def adaptive_avg_pool1d(x: torch.Tensor, t_out: int) -> torch.Tensor:
    wins = adaptive_pool_windows(x.shape[-1], t_out)
    return torch.stack([x[..., s:e].mean(-1) for s, e in wins], dim=-1)


def subject_layers(x, subjects, weights, bias):
    """common.py:45-67 with ``average_subjects=False``: gather per-sample weights, contract over channels."""
    n_subj = weights.shape[0]
    assert subjects.max() < n_subj, "Subject index higher than number of subjects used to initialize the weights."
    idx = subjects.flatten()
    out = torch.einsum("bct,bcd->bdt", x, weights[idx])
    if bias is not None:
        out = out + bias[idx][:, :, None]
    return out


def draw_modality_dropout(modalities: list[str], p: float, training: bool) -> list[str]:
    """model.py:134-141 — one ``torch.rand(1)`` (CPU generator) per modality, *also in eval*; if every modality was
    selected keep one via NumPy's global RNG."""
    dropped = []
    for modality in modalities:
        if torch.rand(1).item() < p and training:
            dropped.append(modality)
    if len(dropped) == len(modalities):
        dropped = list(np.random.choice(dropped, len(dropped) - 1, replace=False))
    return dropped


class OracleFmriEncoder(nn.Module):
    """Parameter creation order follows model.py:46-111 so that a given ``torch.manual_seed`` yields the reference's
    weights: per modality projector (then contrastive head), SubjectLayers weights, bias, time_pos_embed,
    [subject_embed], encoder."""

    def __init__(self, feature_dims, n_outputs, n_output_timesteps, config: OracleConfig,
                 hidden: int = HIDDEN, depth: int = 8, heads: int = 8, xt_semantics: str = "v2"):
        super().__init__()
        self.config, self.feature_dims, self.n_outputs = config, feature_dims, n_outputs
        self.n_output_timesteps, self.hidden = n_output_timesteps, hidden
        self.projectors, self.contrastive_heads = nn.ModuleDict(), nn.ModuleDict()
        n_mod = len(feature_dims)
        for modality, tup in feature_dims.items():
            if tup is None:
                continue
            n_layers, dim = tup
            in_dim = dim * n_layers if config.layer_aggregation == "cat" else dim
            out_dim = hidden // n_mod if config.feature_aggregation == "cat" else hidden
            self.projectors[modality] = nn.Linear(in_dim, out_dim)  # MlpConfig.build shortcut, common.py:124-128
            if config.contrastive_enabled and modality in config.contrastive_modalities:
                self.contrastive_heads[modality] = nn.Linear(in_dim, hidden)
        self.predictor_weights = nn.Parameter(torch.empty(config.n_subjects, hidden, n_outputs).normal_())
        self.predictor_bias = nn.Parameter(torch.empty(config.n_subjects, n_outputs).normal_())
        with torch.no_grad():
            self.predictor_weights *= 1 / hidden**0.5
            self.predictor_bias *= 1 / hidden**0.5
        self.time_pos_embed = nn.Parameter(torch.randn(1, 1024, hidden))
        if config.subject_embedding:
            self.subject_embed = nn.Embedding(config.n_subjects, hidden)
        self.encoder = Encoder(dim=hidden, depth=depth, heads=heads, attn_dim_head=hidden // heads, ff_mult=4,
                               use_scalenorm=True, rotary_pos_emb=True, scale_residual=True, semantics=xt_semantics)

    # -- state-dict bridge to the reference / product naming ----------------------------------------------------
    def reference_state_dict(self):
        sd = {}
        for k, v in self.state_dict().items():
            k = k.replace("predictor_weights", "predictor.weights").replace("predictor_bias", "predictor.bias")
            sd[k] = v
        return sd

    def load_reference_state_dict(self, sd):
        own = {k.replace("predictor.weights", "predictor_weights").replace("predictor.bias", "predictor_bias"): v
               for k, v in sd.items()}
        return self.load_state_dict(own)

    # -- model.py:125-165 -----------------------------------------------------------------------------------------
    def _layer_aggregate(self, data):
        data = data.to(torch.float32)
        if data.ndim == 3:
            data = data.unsqueeze(1)
        if self.config.layer_aggregation == "mean":
            data = data.mean(dim=1)
        else:
            data = data.reshape(data.shape[0], -1, data.shape[-1])  # b l d t -> b (l d) t
        return data.transpose(1, 2)

    def aggregate_features(self, batch):
        ref = next(batch.data[m] for m in batch.data if m in self.feature_dims)
        B, T = ref.shape[0], ref.shape[-1]
        modalities = list(self.feature_dims.keys())
        dropped = draw_modality_dropout(modalities, self.config.modality_dropout, self.training)
        self.last_dropped = list(dropped)
        outs = []
        for modality in modalities:
            if modality not in self.projectors:
                data = torch.zeros(B, T, self.hidden // len(modalities))
            else:
                data = self.projectors[modality](self._layer_aggregate(batch.data[modality]))
                if modality in dropped:
                    data = torch.zeros_like(data)
            outs.append(data)
        return torch.cat(outs, dim=-1) if self.config.feature_aggregation == "cat" else sum(outs)

    def transformer_forward(self, x, subject_id=None):
        x = x + self.time_pos_embed[:, : x.size(1)]
        if hasattr(self, "subject_embed"):
            x = x + self.subject_embed(subject_id)
        return self.encoder(x)

    def forward(self, batch, pool_outputs: bool = True):
        x = self.aggregate_features(batch)
        subject_id = batch.data.get("subject_id", None)
        x = self.transformer_forward(x, subject_id).transpose(1, 2)
        x = subject_layers(x, subject_id, self.predictor_weights, self.predictor_bias)
        return adaptive_avg_pool1d(x, self.n_output_timesteps) if pool_outputs else x

    # -- model.py:177-241 -----------------------------------------------------------------------------------------
    @staticmethod
    def info_nce(q, k, tau=0.07):
        q = F.normalize(q.reshape(-1, q.shape[-1]), dim=-1)
        k = F.normalize(k.reshape(-1, k.shape[-1]), dim=-1)
        logits = q @ k.t() / tau
        labels = torch.arange(logits.shape[0])
        return 0.5 * (F.cross_entropy(logits, labels) + F.cross_entropy(logits.t(), labels))

    def compute_contrastive_loss(self, batch):
        if not self.config.contrastive_enabled:
            return {}
        brain = self.transformer_forward(self.aggregate_features(batch), batch.data.get("subject_id", None))
        losses = {}
        for modality in self.config.contrastive_modalities:
            if modality not in self.contrastive_heads or modality not in batch.data:
                continue
            lat = self.contrastive_heads[modality](self._layer_aggregate(batch.data[modality]))
            if lat.size(1) != brain.size(1):
                lat = adaptive_avg_pool1d(lat.transpose(1, 2), brain.size(1)).transpose(1, 2)
            losses[modality] = self.info_nce(brain, lat, self.config.contrastive_temperature)
        return losses


def flatten_bdt(x):
    """``rearrange(x, "b d t -> (b t) d")`` (pl_module.py:54-55)."""
    return x.permute(0, 2, 1).reshape(-1, x.shape[1])


def run_step(model: OracleFmriEncoder, batch, loss_fn=None):
    """pl_module.py:46-107 minus logging: returns (loss, y_pred, y_true, contrastive dict)."""
    loss_fn = loss_fn or nn.MSELoss()
    y_true = batch.data["fmri"]
    y_pred = model(batch)
    loss = loss_fn(flatten_bdt(y_pred), flatten_bdt(y_true))
    contrastive = model.compute_contrastive_loss(batch)
    if contrastive:
        total = sum(contrastive.values()) / max(1, len(contrastive))
        loss = loss + model.config.contrastive_weight * total
    return loss, y_pred, y_true, contrastive


def pearson_loss(x, y, reduction="mean", dim=1):
    """losses.py:11-42."""
    x, y = x.transpose(0, dim), y.transpose(0, dim)
    x, y = x.reshape(x.shape[0], -1), y.reshape(y.shape[0], -1)
    x, y = x - x.mean(1, keepdim=True), y - y.mean(1, keepdim=True)
    pcc = (x * y).sum(1) / ((x.pow(2).sum(1).sqrt() * y.pow(2).sum(1).sqrt()) + 1e-8)
    loss = 1 - pcc
    return loss.mean() if reduction == "mean" else loss.sum()


def multidim_pearson_scipy(preds: np.ndarray, trues: np.ndarray) -> np.ndarray:
    """main.py:470-477: ``(b d t) -> (b t) d`` then a Python loop of ``scipy.stats.pearsonr`` per parcel."""
    from scipy.stats import pearsonr

    preds = np.ascontiguousarray(np.transpose(preds, (0, 2, 1))).reshape(-1, preds.shape[1])
    trues = np.ascontiguousarray(np.transpose(trues, (0, 2, 1))).reshape(-1, trues.shape[1])
    out = np.zeros(trues.shape[1], dtype=np.float32)
    for p in range(len(out)):
        out[p] = pearsonr(trues[:, p], preds[:, p])[0]
    return out


def pearson_columns_f64(preds: np.ndarray, trues: np.ndarray) -> np.ndarray:
    """float64 two-pass per-column Pearson on (N, O) matrices (independent cross-check of scipy)."""
    x, y = preds.astype(np.float64), trues.astype(np.float64)
    x, y = x - x.mean(0), y - y.mean(0)
    return (x * y).sum(0) / np.sqrt((x * x).sum(0) * (y * y).sum(0))


def streaming_pearson(pred_batches, true_batches, num_outputs):
    """MultidimPearsonCorrCoef (metrics/base.py:26-29) over a sequence of (n, O) batches: per-parcel r and mean."""
    m = PearsonCorrCoef(num_outputs=num_outputs)
    for p, t in zip(pred_batches, true_batches):
        m.update(p, t)
    r = m.compute()
    return r, r.mean()


def ensemble_average(member_preds: np.ndarray, member_pearson: np.ndarray, temperature: float = 0.3) -> np.ndarray:
    """average_submissions.py:107-125: per-parcel softmax(pearson / tau) weights over members, weighted sum of
    predictions.  member_preds (N, n_tr, O), member_pearson (N, O) -> (n_tr, O)."""
    w = member_pearson / temperature
    w = np.exp(w - w.max(0, keepdims=True))
    w = w / w.sum(0, keepdims=True)
    return (member_preds * w[:, None, :]).sum(0)


def synthetic_batch(batch_size=2, t=298, t_out=100, n_outputs=1000, n_subjects=4, seed=1234,
                    dims=(("text", 2, 3072), ("audio", 2, 1024), ("video", 2, 1408)), dtype=torch.float32):
    """SURVEY §8(d) synthetic window batch: N(0,1) features, N(0,1) fMRI, uniform subject ids."""
    g = torch.Generator().manual_seed(seed)
    data = {name: torch.randn(batch_size, l, d, t, generator=g).to(dtype) for name, l, d in dims}
    data["fmri"] = torch.randn(batch_size, n_outputs, t_out, generator=g)
    data["subject_id"] = torch.randint(0, n_subjects, (batch_size, 1), generator=g)
    return SegmentData(data=data, segments=[None] * batch_size)


def train_flops_per_window(t=298, hidden=HIDDEN, depth=8, n_out=1000, k_proj=11008, contrastive=False):
    """SURVEY §8(d) algorithmic FLOPs (forward x3 for a train step)."""
    proj = 2 * t * k_proj * (hidden // 3)
    enc = depth * (4 * 2 * t * hidden**2 + 2 * 2 * t * hidden * 4 * hidden)
    attn = depth * 2 * (2 * t * t * hidden)
    readout = 2 * t * hidden * n_out
    fwd = proj + enc + attn + readout
    return 3 * fwd if not contrastive else 3 * (fwd + proj + enc + attn)


assert math.isclose(train_flops_per_window() / 1e12, 1.672, rel_tol=2e-3)
