"""``FmriEncoder`` / ``FmriEncoderConfig`` — drop-in mirror of the reference's ``algonauts2025/model.py`` whose math
runs on the sm_100a kernels (``engine.py``).  Same constructor, method names, attributes, parameter names (so
``state_dict`` round-trips with reference checkpoints), same CPU-RNG consumption for modality dropout, same error
behaviour; no CPU fallback (the module owns a CUDA device).

Reference anchors: ``FmriEncoderConfig`` model.py:20-43; ``FmriEncoder.__init__`` :46-111; ``forward`` :113-123;
``aggregate_features`` :125-165; ``transformer_forward`` :167-174; contrastive helpers :177-241;
``SubjectLayers`` modeling_utils/modeling_utils/models/common.py:14-71; encoder = x_transformers ``Encoder`` built at
modeling_utils/modeling_utils/models/transformer.py:43-61 (semantics declared in DESIGN.md).
"""
from __future__ import annotations

import typing as tp

import numpy as np
import pydantic
import torch
from torch import nn

from . import ops
from ._lib import TribeError
from .engine import Engine, Plan
from .segment import SegmentData

HIDDEN = 3072  # model.py:61


class FmriEncoderConfig(pydantic.BaseModel):
    model_config = pydantic.ConfigDict(extra="forbid")
    name: tp.Literal["FmriEncoder"] = "FmriEncoder"
    n_subjects: int | None = None
    feature_aggregation: tp.Literal["sum", "cat"] = "cat"
    layer_aggregation: tp.Literal["mean", "cat"] = "cat"
    subject_embedding: bool = False
    modality_dropout: float = 0.0

    # Contrastive alignment (e.g., with VJEPA2 video features)
    contrastive_enabled: bool = False
    contrastive_modalities: list[str] = ["video"]
    contrastive_weight: float = 0.1
    contrastive_temperature: float = 0.07
    # extension (not in the reference schema, hence excluded from model_dump() — experiment ids / config hashes stay the
    # reference's): arithmetic of the third-party encoder, see FmriEncoder.__init__
    xt_semantics: tp.Literal["v2", "v1.27"] = pydantic.Field("v2", exclude=True)

    def build(self, feature_dims: dict, n_outputs: int, n_output_timesteps: int) -> nn.Module:
        return FmriEncoder(feature_dims, n_outputs, n_output_timesteps, config=self)


# ---------------------------------------------------------------------------------------------------- containers
class _Fn(torch.autograd.Function):
    """Glue between torch autograd (``loss.backward()`` under Lightning's automatic optimisation) and the engine."""

    @staticmethod
    def forward(ctx, anchor, x_in, engine, plan, batch_data):
        out = engine.forward(plan, batch_data, x_in)
        ctx.engine, ctx.plan = engine, plan
        ctx.want_dx = x_in is not None and x_in.requires_grad
        return out

    @staticmethod
    def backward(ctx, grad_out):
        dx = ctx.engine.backward(ctx.plan, grad_out, want_dx_in=ctx.want_dx)
        return None, dx, None, None, None


class _SubjectLayersFn(torch.autograd.Function):
    """Stand-alone ``SubjectLayers.forward`` with its backward on the same tcgen05 GEMMs the engine uses for the readout."""

    @staticmethod
    def forward(ctx, x, weights, bias, subj):
        B, C, T = x.shape
        N, _, D = weights.shape
        dev = x.device
        xt = torch.empty(B * T, C, device=dev, dtype=torch.bfloat16)
        ops.ingest_features(x.detach().unsqueeze(1), xt, 0, False)  # (B, C, T) -> bf16 (B*T, C)
        w16 = weights.detach().to(torch.bfloat16).contiguous()
        out = torch.empty(B, D, T, device=dev, dtype=torch.float32)
        x_op = ops.Operand(xt, inner=C, rows=T, row_stride=C, batch=B, batch_stride=T * C)
        w_op = ops.Operand(w16, inner=D, rows=C, row_stride=D, batch=N, batch_stride=C * D, mn_major=True, gather=subj)
        b32 = bias.detach().float().contiguous() if bias is not None else None
        ops.gemm(x_op, w_op, out, T, D, C, ldd=T, batch=B, d_zo=D * T, transposed=True, bias=b32, bias_gathered=b32 is not None,
                 bias_z_stride=D)
        ctx.save_for_backward(xt, w16, subj)
        ctx.dims, ctx.has_bias = (B, C, T, N, D), bias is not None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        xt, w16, subj = ctx.saved_tensors
        B, C, T, N, D = ctx.dims
        dev = grad_out.device
        dyT = torch.empty(B, T, D, device=dev, dtype=torch.bfloat16)
        ops.transpose_cast_bot(grad_out.contiguous().float(), dyT)
        dx = dW = dB = None
        if ctx.needs_input_grad[0]:
            # dx[b] (C, T) = W[s_b] (C, D) dy[b] (D, T): per sample (T, D) x (C, D)^T, stored transposed
            dx = torch.empty(B, C, T, device=dev, dtype=torch.float32)
            dy_op = ops.Operand(dyT, inner=D, rows=T, row_stride=D, batch=B, batch_stride=T * D)
            w_op = ops.Operand(w16, inner=D, rows=C, row_stride=D, batch=N, batch_stride=C * D, gather=subj)
            ops.gemm(dy_op, w_op, dx, T, C, D, ldd=T, batch=B, d_zo=C * T, transposed=True)
        if ctx.needs_input_grad[1]:
            # dW[s] (C, D) = sum_{b: s_b = s} x[b] (C, T) dy[b]^T (T, D): subject-keyed K loop, no per-sample (B, C, D) tensor
            dW = torch.empty(N, C, D, device=dev, dtype=torch.float32)
            xa = ops.Operand(xt, inner=C, rows=T, row_stride=C, batch=B, batch_stride=T * C, mn_major=True)
            dyb = ops.Operand(dyT, inner=D, rows=T, row_stride=D, batch=B, batch_stride=T * D, mn_major=True)
            ops.gemm(xa, dyb, dW.view(N * C, D), C, D, T, ldd=D, batch=N, d_zo=C * D, kgroup=subj)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            dB = torch.empty(N, D, device=dev, dtype=torch.float32)
            ops.zero_(dB)
            ops.subject_bias_grad(dyT, subj, dB, B, T, D, N)
        return dx, dW, dB, None


class SubjectLayers(nn.Module):
    """Per-subject linear readout (common.py:14-71).  ``forward(x (B, C, T), subjects (B, 1))`` runs the
    subject-index-gathered grouped GEMM; inside ``FmriEncoder`` the engine calls the same GEMM on pooled tokens."""

    def __init__(self, in_channels: int, out_channels: int, n_subjects: int, bias: bool = False, init_id: bool = False,
                 average_subjects: bool = False):
        super().__init__()
        self.weights = nn.Parameter(torch.empty(n_subjects, in_channels, out_channels))
        self.bias = nn.Parameter(torch.empty(n_subjects, out_channels)) if bias else None
        if init_id:
            if in_channels != out_channels:
                raise ValueError("in_channels and out_channels must be the same for identity initialization.")
            self.weights.data[:] = torch.eye(in_channels)[None]
            if self.bias is not None:
                self.bias.data[:] = 0
        else:
            self.weights.data.normal_()
            if self.bias is not None:
                self.bias.data.normal_()
        self.weights.data *= 1 / in_channels**0.5
        if self.bias is not None:
            self.bias.data *= 1 / in_channels**0.5
        if average_subjects:
            raise NotImplementedError("average_subjects=True is not on the TRIBE path (model.py:99)")
        self.average_subjects = average_subjects

    def forward(self, x: torch.Tensor, subjects: torch.Tensor) -> torch.Tensor:
        """common.py:45-67 stand-alone: ``einsum("bct,bcd->bdt", x, weights[subjects]) + bias[subjects]``, differentiable
        w.r.t. ``x``, ``weights`` and ``bias`` (gathered dgrad GEMM, subject-keyed grouped wgrad, bias reduction)."""
        if not x.is_cuda:
            raise TribeError("SubjectLayers needs CUDA tensors (no CPU fallback)")
        if x.dim() != 3 or x.shape[1] != self.weights.shape[1]:
            raise TribeError(f"SubjectLayers expects x (B, {self.weights.shape[1]}, T), got {tuple(x.shape)}")
        if subjects.numel() != x.shape[0]:
            raise TribeError(f"SubjectLayers: {subjects.numel()} subject ids for a batch of {x.shape[0]}")
        N = self.weights.shape[0]
        assert subjects.max() < N, "Subject index higher than number of subjects used to initialize the weights."
        subj = subjects.flatten().to(x.device, torch.int64).contiguous()
        return _SubjectLayersFn.apply(x, self.weights.to(x.device), self.bias.to(x.device) if self.bias is not None else None, subj)

    def __repr__(self):
        S, C, D = self.weights.shape
        return f"SubjectLayers({C}, {D}, {S})"


class _PoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, t_out):
        ctx.t_in = x.shape[-1]
        return ops.adaptive_avg_pool_fwd(x.float(), t_out)

    @staticmethod
    def backward(ctx, dy):
        return ops.adaptive_avg_pool_bwd(dy.float(), ctx.t_in), None


class AdaptiveAvgPool1d(nn.Module):
    """``nn.AdaptiveAvgPool1d(n_output_timesteps)`` (model.py:60) as the coalesced bandwidth kernel."""

    def __init__(self, output_size: int):
        super().__init__()
        self.output_size = output_size

    def forward(self, x):
        if not x.is_cuda:
            raise TribeError("AdaptiveAvgPool1d needs CUDA tensors (no CPU fallback)")
        return _PoolFn.apply(x, self.output_size)

    def extra_repr(self):
        return f"output_size={self.output_size}"


class _ScaleNorm(nn.Module):
    def __init__(self, g_init: float = 1.0):
        super().__init__()
        self.g = nn.Parameter(torch.ones(1) * g_init)


class _Attention(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(dim, dim, bias=False)
        self.to_v = nn.Linear(dim, dim, bias=False)
        self.to_out = nn.Linear(dim, dim, bias=False)


class _FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.ff = nn.Sequential(nn.Sequential(nn.Linear(dim, dim * mult), nn.GELU()), nn.Dropout(0.0), nn.Linear(dim * mult, dim))


class _Residual(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.residual_scale = nn.Parameter(torch.ones(dim))


class _Rotary(nn.Module):
    def __init__(self, dim, base=10000.0):
        super().__init__()
        self.register_buffer("inv_freq", 1.0 / (base ** (torch.arange(0, dim, 2).float() / dim)))


class TribeEncoder(nn.Module):
    """Parameter container with x_transformers' module nesting (``layers.<i>.0.0.g``, ``layers.<i>.1.to_q.weight``,
    ``layers.<i>.1.ff.0.0.weight``, ``layers.<i>.2.residual_scale``, ``final_norm.g``, ``rotary_pos_emb.inv_freq``);
    creation order = x_transformers' so a seed reproduces the reference's init.  ``forward`` runs the fused engine."""

    def __init__(self, dim, depth, heads, semantics: str = "v2"):
        super().__init__()
        self.dim, self.depth, self.heads, self.semantics = dim, depth, heads, semantics
        g_init = dim ** -0.5 if semantics == "v1.27" else 1.0  # 1.27.x ScaleNorm: g = ones(1) * dim ** -0.5, no sqrt(dim) factor
        self.rotary_pos_emb = _Rotary(max(dim // heads // 2, 32))
        layers = []
        for _ in range(depth):
            for kind in ("a", "f"):
                block = _Attention(dim) if kind == "a" else _FeedForward(dim)
                layers.append(nn.ModuleList([nn.ModuleList([_ScaleNorm(g_init), None, None]), block, _Residual(dim)]))
        self.layers = nn.ModuleList(layers)
        self.final_norm = _ScaleNorm(g_init)
        self._owner = None

    def forward(self, x):
        owner = self._owner() if self._owner is not None else None
        if owner is None:
            raise TribeError("TribeEncoder must be used through its FmriEncoder")
        return owner._encode_tensor(x, add_pos=False)


class FmriEncoder(nn.Module):
    def __init__(self, feature_dims: dict[str, tuple[int, int]], n_outputs: int, n_output_timesteps: int, config: FmriEncoderConfig,
                 *, hidden: int = HIDDEN, depth: int = 8, heads: int = 8, device=None, xt_semantics: str | None = None):
        """``xt_semantics`` (extension; also a config field): which x_transformers release's arithmetic the fusion encoder
        follows — ``"v2"`` (>= 1.30 / 2.x, default) or ``"v1.27"`` (the oldest release modeling_utils/pyproject.toml:12
        admits).  State-dict keys are the same under both, so a checkpoint must be evaluated with the semantics of the
        library version that trained it."""
        super().__init__()
        self.xt_semantics = xt_semantics or getattr(config, "xt_semantics", "v2")
        if self.xt_semantics not in ("v2", "v1.27"):
            raise ValueError(f"xt_semantics must be 'v2' or 'v1.27', got {self.xt_semantics!r}")
        self.config = config
        self.feature_dims = feature_dims
        self.n_outputs = n_outputs
        self.hidden, self.depth, self.heads = hidden, depth, heads
        self.projectors = nn.ModuleDict()
        self.contrastive_heads = nn.ModuleDict()
        self.pooler = AdaptiveAvgPool1d(n_output_timesteps)
        for modality, tup in feature_dims.items():
            if tup is None:
                print(f"Warning: {modality} has no feature dimensions. Skipping projector.")
                continue
            num_layers, feature_dim = tup
            input_dim = feature_dim * num_layers if config.layer_aggregation == "cat" else feature_dim
            output_dim = hidden // len(feature_dims) if config.feature_aggregation == "cat" else hidden
            # MlpConfig(...).build(in, out) returns a bare nn.Linear when hidden_sizes is None (common.py:124-128)
            self.projectors[modality] = nn.Linear(input_dim, output_dim)
            if config.contrastive_enabled and modality in config.contrastive_modalities:
                self.contrastive_heads[modality] = nn.Linear(input_dim, hidden)
        self.combiner = nn.Identity()
        self.predictor = SubjectLayers(in_channels=hidden, out_channels=n_outputs, n_subjects=config.n_subjects,
                                       average_subjects=False, bias=True)
        self.time_pos_embed = nn.Parameter(torch.randn(1, 1024, hidden))
        if config.subject_embedding:
            self.subject_embed = nn.Embedding(config.n_subjects, hidden)
        if hidden % heads != 0:
            raise ValueError(f"dim ({hidden}) must be divisible by the number of heads ({heads})")  # transformer.py:46-49
        if hidden < 256:
            raise ValueError(f"dim ({hidden}) is less than 256, which causes weird bug in x-transformers")  # :50-53
        self.encoder = TribeEncoder(hidden, depth, heads, self.xt_semantics)
        import weakref

        self.encoder._owner = weakref.ref(self)
        self._engine = Engine(self)
        if device is None and torch.cuda.is_available():
            device = torch.device("cuda", torch.cuda.current_device())
        if device is not None:
            self._engine.materialize(device)

    # -- device ownership: parameters live in one flat CUDA buffer; nn.Module.to()/cuda() keep working ---------------
    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        eng = self.__dict__.get("_engine")
        if eng is not None and eng.flat is not None and not eng.flat.intact():
            dev = next((p.device for p in self.parameters() if p.is_cuda), None)
            eng.flat = None
            if dev is not None:
                eng.materialize(dev)
        return out

    @property
    def device(self):
        return self._engine.device

    def state_dict(self, *args, **kwargs):
        flat = self._engine.flat
        if flat is not None and flat.sharded is not None:
            flat.sharded.gather_masters()  # collective: rank-sharded optimizer keeps fp32 masters on their owner
        return super().state_dict(*args, **kwargs)

    # -- forward paths ----------------------------------------------------------------------------------------------
    def _draw_dropout(self) -> list[str]:
        """model.py:134-141 verbatim in behaviour: one CPU ``torch.rand(1)`` per modality *evaluated before*
        ``and self.training`` (so eval also advances the generator); NumPy global RNG if everything was selected."""
        preset = self.__dict__.get("_preset_dropped")
        if preset:  # drawn ahead of a CUDA-graph capture / replay by graphed.GraphedTrainStep (same draws, same order)
            return list(preset.pop(0))
        dropped = []
        for modality in self.feature_dims.keys():
            if torch.rand(1).item() < self.config.modality_dropout and self.training:
                dropped.append(modality)
        if len(dropped) == len(self.feature_dims):
            dropped = list(np.random.choice(dropped, len(dropped) - 1, replace=False))
        return dropped

    # common.py:53-55 asserts ``subjects.max() < N`` on the host (a device sync in the reference as well).  The range
    # check always runs on the device into a sticky per-model flag.  With ``defer_subject_check = True`` (set by
    # trainer.MiniTrainer) the flag is copied to pinned host memory asynchronously and examined at the NEXT call /
    # ``flush_subject_check()``, so the training loop has no host-device synchronisation point and the whole step can
    # be captured in a CUDA graph (the copy is a memcpy node).
    defer_subject_check = False
    _MSG = "Subject index higher than number of subjects used to initialize the weights."

    def flush_subject_check(self) -> None:
        flags = self.__dict__.get("_subject_flags")
        if flags is not None:
            torch.cuda.current_stream(flags[0].device).synchronize()
            ops.drain_stale_error("before reading the subject-range flag")
            bad = int(flags[0].item())
            flags[0].zero_()
            flags[1].zero_()
            self.__dict__["_subject_event"] = None
            assert bad == 0, self._MSG

    def _subjects(self, batch, subject_id=None):
        if subject_id is None:
            subject_id = batch.data.get("subject_id", None) if batch is not None else None
        if subject_id is None:
            return None
        self._engine._check_flat()
        dev = self._engine.device
        subj = subject_id.to(dev, torch.int64, non_blocking=True).flatten().contiguous()
        n = self.predictor.weights.shape[0]
        flags = self.__dict__.get("_subject_flags")
        if flags is None or flags[0].device != dev:
            flags = self.__dict__["_subject_flags"] = (torch.zeros(1, device=dev, dtype=torch.int32), torch.zeros(1, dtype=torch.int32).pin_memory())
        dev_flag, host_flag = flags
        if self.defer_subject_check:
            pending = self.__dict__.get("_subject_event")
            if pending is not None:
                pending.synchronize()  # the PREVIOUS call's flag copy: the host runs at most one step ahead
            if int(host_flag.item()) != 0:
                self.flush_subject_check()
            safe = torch.empty_like(subj)  # ids clamped into range: the enqueued step stays in bounds until the flag is read
            ops.check_subjects(subj, n, dev_flag, safe)
            subj = safe
            ops.copy_to_pinned(host_flag, dev_flag)  # (not Tensor.copy_: see ops.copy_to_pinned)
            self.__dict__["_subject_event"] = None
            if not torch.cuda.is_current_stream_capturing():  # graph replays poll host_flag themselves (graphed.py)
                event = torch.cuda.Event()
                event.record()
                self.__dict__["_subject_event"] = event
        else:
            ops.check_subjects(subj, n, dev_flag)
            self.flush_subject_check()
        return subj

    # Gradients are written by hand into the flat buffer (engine.backward): the parameters' AccumulateGrad nodes never run,
    # so torch DistributedDataParallel's reducer hooks never fire — wrapping this module in DDP (Lightning strategy "ddp*",
    # main.py:388-394) would silently train every rank on its local gradient.  Data-parallel training goes through
    # parallel.data_parallel(model, optimizer) (BrainModule.configure_optimizers installs it); a training forward in a
    # multi-rank job without it raises unless the replica is declared independent (ensemble members).
    independent_replica = False

    def _check_gradient_sync(self) -> None:
        if not self.training or self._engine.comm is not None or self.independent_replica:
            return
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            raise TribeError("training forward in a multi-rank job without gradient synchronisation: torch DDP hooks never fire for this "
                             "module (gradients are written directly into the flat buffer). Use parallel.data_parallel(model, optimizer) "
                             "(BrainModule.configure_optimizers does) or set model.independent_replica = True for ensemble members.")

    def _anchor(self):
        # What makes the output of ``_Fn`` require grad.  A FRESH leaf per call (not a parameter): autograd caches a
        # leaf's AccumulateGrad node together with the stream it was created on for as long as any old graph is alive,
        # and running that node during a CUDA-graph capture would tie the capture to uncaptured work on that stream.
        return torch.empty(0, device=self._engine.device, requires_grad=True)

    def _run(self, batch, *, mode, pool=True, x_in=None, subject_id=None, out=None):
        eng = self._engine
        eng._check_flat()
        plan = Plan()
        plan.out = out
        plan.mode, plan.pool, plan.t_out = mode, pool, self.pooler.output_size
        plan.training = self.training
        if x_in is None:
            plan.dropped = self._draw_dropout()
            self.last_dropped = list(plan.dropped)
            plan.subjects = self._subjects(batch) if (mode == "predict" or hasattr(self, "subject_embed")) else None
            data = batch.data
        else:
            plan.x_input = True
            data = None
            if hasattr(self, "subject_embed"):
                if subject_id is None:
                    raise TribeError("transformer_forward: subject_embedding=True needs subject_id (model.py:171-172)")
                plan.subjects = self._subjects(None, subject_id)
                if plan.subjects.numel() != x_in.shape[0]:
                    raise TribeError(f"transformer_forward: {plan.subjects.numel()} subject ids for a batch of {x_in.shape[0]}")
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        plan.keep = needs_grad
        if needs_grad:
            self._check_gradient_sync()
            out = _Fn.apply(self._anchor(), x_in, eng, plan, data)
            plan.awaiting = eng  # counted until its backward runs or autograd drops the graph (engine.Plan.settle)
            eng._grad_plans += 1
            return out
        out = eng.forward(plan, data, x_in)
        eng.release(plan)
        return out

    def forward(self, batch: SegmentData, pool_outputs: bool = True, out: torch.Tensor | None = None) -> torch.Tensor:
        """model.py:113-123.  ``out`` (extension): a preallocated (B, n_outputs, T') fp32 CUDA tensor the readout GEMM writes
        into — evaluation sweeps collect predictions without a copy (no-grad calls only)."""
        if out is not None and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise TribeError("forward(out=...) is for inference: call it under torch.no_grad()")
        return self._run(batch, mode="predict", pool=pool_outputs, out=out)

    def aggregate_features(self, batch):
        """model.py:125-165 (projector outputs, concatenated / summed; dropped modalities zeroed), (B, T, H) fp32.
        Stand-alone entry (main.py:346 logs its shape); inside ``forward`` the same GEMMs additionally fuse the
        positional embedding."""
        eng = self._engine
        eng._check_flat()
        eng.flat.refresh_bf16()
        dropped = self._draw_dropout()
        self.last_dropped = list(dropped)
        B, T = eng._validate_batch(batch.data)
        M, H = B * T, self.hidden
        mods = list(self.feature_dims.keys())
        cat = self.config.feature_aggregation == "cat"
        width = H // len(mods) if cat else H
        out = torch.zeros(M, H, device=eng.device, dtype=torch.float32)
        first = True
        for i, mod in enumerate(mods):
            if mod not in self.projectors or mod in dropped:
                continue
            col = i * width if cat else 0
            K = self.projectors[mod].in_features
            feat = torch.empty(M, K, device=eng.device, dtype=torch.bfloat16)
            ops.ingest_features(batch.data[mod].to(eng.device), feat, 0, self.config.layer_aggregation == "mean")
            w16, bias = eng._w16(f"projectors.{mod}.weight"), eng._p(f"projectors.{mod}.bias")
            if cat or first:
                ops.gemm(ops.kmajor(feat), ops.kmajor(w16), out, M, width, K, ldd=H, d_off=col, bias=bias)
            else:
                ops.gemm(ops.kmajor(feat), ops.kmajor(w16), out, M, width, K, ldd=H, bias=bias, epilogue=ops.EPI_RESIDUAL, res=out, ld_res=H)
            first = False
        return out.view(B, T, H)

    def _encode_tensor(self, x, add_pos=True, subject_id=None):
        if not add_pos:
            raise TribeError("calling .encoder(x) directly is not supported; use transformer_forward")
        if x.dim() != 3 or x.shape[-1] != self.hidden:
            raise TribeError(f"transformer_forward expects (B, T, {self.hidden}), got {tuple(x.shape)}")
        return self._run(None, mode="latents", x_in=x, subject_id=subject_id)

    def transformer_forward(self, x, subject_id=None):
        """model.py:167-174: ``x + time_pos_embed[:, :T]`` (+ ``subject_embed(subject_id)`` broadcast over T when the
        model was built with ``subject_embedding=True``) -> encoder.  Differentiable w.r.t. ``x`` and every parameter."""
        return self._encode_tensor(x, add_pos=True, subject_id=subject_id)

    # --- Contrastive alignment helpers (model.py:177-241) -------------------------------------------------------------
    def get_brain_latents(self, batch: SegmentData) -> torch.Tensor:
        """Sequence latents before the predictor, (B, T, H): a second projector + encoder pass with fresh dropout
        draws, exactly like the reference (model.py:178-183)."""
        return self._run(batch, mode="latents")

    def get_modality_latents(self, batch: SegmentData, modality: str) -> torch.Tensor:
        from .contrastive import modality_latents

        assert modality in self.contrastive_heads, f"No contrastive head found for modality '{modality}'"
        if batch.data.get(modality, None) is None:
            raise KeyError(f"Modality '{modality}' not found in batch.data")
        return modality_latents(self, batch, modality)

    @staticmethod
    def _info_nce(q: torch.Tensor, k: torch.Tensor, tau: float = 0.07) -> torch.Tensor:
        from .contrastive import info_nce

        return info_nce(q, k, tau)

    def compute_contrastive_loss(self, batch: SegmentData) -> dict[str, torch.Tensor]:
        if not self.config.contrastive_enabled:
            return {}
        tau = self.config.contrastive_temperature
        brain_latents = self.get_brain_latents(batch)
        losses: dict[str, torch.Tensor] = {}
        for modality in self.config.contrastive_modalities:
            if modality not in self.contrastive_heads or modality not in batch.data:
                continue
            mod_latents = self.get_modality_latents(batch, modality)
            if mod_latents.size(1) != brain_latents.size(1):
                mod_latents = self.pooler.__class__(brain_latents.size(1))(mod_latents.transpose(1, 2).contiguous()).transpose(1, 2)
            losses[modality] = self._info_nce(brain_latents, mod_latents, tau=tau)
        return losses
