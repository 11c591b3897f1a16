"""Forward / backward execution of the TRIBE encoding model on the sm_100a kernels (``ops``).

This is the B200 replacement for everything ``FmriEncoder.forward`` triggers in the reference
(``algonauts2025/model.py:113-174``): feature ingest -> projector GEMMs (+bias +positional embedding fused) ->
8 x [ScaleNorm -> fused QKV GEMM (+rotary) -> batched QK^T -> fp32 softmax -> batched PV -> out-proj GEMM
(+scaled residual) ; ScaleNorm -> FF1 GEMM (+bias +GELU) -> FF2 GEMM (+bias +scaled residual)] -> final ScaleNorm ->
token pooling -> subject-gathered readout GEMM, and the hand-written backward of the same graph.

Numerics: bf16 tensor-core operands, fp32 accumulation, fp32 residual stream / softmax / norm statistics, fp32 master
weights and gradients.  All buffers are torch allocations; all math is in ``libtribe_b200.so``.
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import TribeError

ALIGN = 64  # elements; keeps every parameter view 16-byte aligned in fp32 and bf16

# Fused CUDA optimizers update parameters through raw pointers, so the per-tensor version counters alone are not a
# reliable "weights changed" signal for the bf16 shadow copy: a global optimizer post-step hook bumps a counter on every
# FlatParams whose parameters the stepping optimizer owns.
_FLATS = []  # weak references to live FlatParams
_HOOK = []


def _install_optimizer_hook():
    if not _HOOK:
        from torch.optim.optimizer import register_optimizer_step_post_hook

        def _bump(optimizer, args, kwargs):
            alive = []
            for ref in _FLATS:
                flat = ref()
                if flat is None:
                    continue
                alive.append(ref)
                if any(id(p) in flat.param_ids for group in optimizer.param_groups for p in group["params"][:1]):
                    flat.opt_steps += 1
            _FLATS[:] = alive

        _HOOK.append(register_optimizer_step_post_hook(_bump))


def _round_up(x: int, a: int) -> int:
    return (x + a - 1) // a * a


class FlatParams:
    """All parameters of the model as views of ONE flat fp32 buffer (+ a flat fp32 gradient buffer and a flat bf16
    shadow copy at the same offsets): one cast kernel refreshes every tensor-core operand after an optimizer step and
    gradient buckets for the data-parallel all-reduce are plain contiguous slices."""

    def __init__(self, named_params, buckets, device):
        self.device = device
        self.names = [n for n, _ in named_params]
        self.params = dict(named_params)
        order = [n for b in buckets for n in b]
        assert sorted(order) == sorted(self.names), "bucket plan must cover every parameter exactly once"
        self.offsets, off = {}, 0
        self.bucket_ranges = []
        for b in buckets:
            start = off
            for n in b:
                self.offsets[n] = off
                off += _round_up(self.params[n].numel(), ALIGN)
            self.bucket_ranges.append((start, off))
        self.total = off
        self.flat = torch.zeros(self.total, device=device, dtype=torch.float32)
        for n, p in self.params.items():
            view = self.flat[self.offsets[n]: self.offsets[n] + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
        self.grad = None
        self.bf16 = None
        self._sig = None
        self.param_ids = {id(p) for p in self.params.values()}
        self.opt_steps = 0
        self.sharded = None  # parallel.ShardedStep: fp32 masters / Adam moments are only current on their owner rank
        import weakref

        _FLATS.append(weakref.ref(self))
        _install_optimizer_hook()

    def intact(self) -> bool:
        base = self.flat.data_ptr()
        return all(p.data_ptr() == base + 4 * self.offsets[n] for n, p in self.params.items())

    def view16(self, name):
        p = self.params[name]
        o = self.offsets[name]
        return self.bf16[o: o + p.numel()].view(p.shape)

    def gview(self, name):
        p = self.params[name]
        o = self.offsets[name]
        return self.grad[o: o + p.numel()].view(p.shape)

    def ensure_grad(self):
        if self.grad is None:
            self.grad = torch.zeros(self.total, device=self.device, dtype=torch.float32)
        return self.grad

    def refresh_bf16(self):
        """Re-cast the bf16 shadow when any parameter changed in place (optimizer step, load_state_dict, SWA)."""
        sig = (sum(p._version for p in self.params.values()), self.opt_steps)
        if self.bf16 is None:
            self.bf16 = torch.empty(self.total, device=self.device, dtype=torch.bfloat16)
            self._sig = None
        if sig != self._sig:
            if self.sharded is not None:
                # rank-sharded optimizer: a foreign in-place change (load_state_dict, SWA swap — the same calls on every
                # rank) must not be cast from masters that are stale outside this rank's slices
                self.sharded.gather_masters()
            ops.cast_f32_bf16(self.flat, self.bf16)
            self._sig = sig

    def rehome(self, alloc):
        """Move the flat fp32 masters, the gradient buffer and the bf16 shadow into buffers from ``alloc(numel, dtype)``
        (symmetric memory for the NVLink step tail, parallel.ShardedStep); every parameter / gradient view follows."""
        flat, grad, bf16 = alloc(self.total, torch.float32), alloc(self.total, torch.float32), alloc(self.total, torch.bfloat16)
        flat.copy_(self.flat)
        grad.zero_()
        if self.bf16 is not None:
            bf16.copy_(self.bf16)
        else:
            ops.cast_f32_bf16(flat, bf16)
            self._sig = (sum(p._version for p in self.params.values()), self.opt_steps)
        versions = sum(p._version for p in self.params.values())
        self.flat, self.grad, self.bf16 = flat, grad, bf16
        for n, p in self.params.items():
            p.data = self.flat[self.offsets[n]: self.offsets[n] + p.numel()].view(p.shape)
            p.grad = None
        if self._sig is not None:  # re-pointing .data must not look like a weight change
            self._sig = (self._sig[0] + sum(p._version for p in self.params.values()) - versions, self._sig[1])


class Workspace:
    """Activations + scratch for one (batch, T) shape.  A training forward keeps its workspace until its backward ran
    (the contrastive branch runs the encoder twice per step, pl_module.py:59-77), so workspaces are pooled."""

    def __init__(self, eng, B, T):
        dev, H, F = eng.device, eng.hidden, eng.ff
        M = B * T
        self.B, self.T, self.M = B, T, M
        self.Tp = _round_up(T, 8)
        nsub = 2 * eng.depth
        bf, f32 = torch.bfloat16, torch.float32
        e = lambda *s, dtype=bf: torch.empty(*s, device=dev, dtype=dtype)  # noqa: E731
        self.feat = {m: e(M, k) for m, k in eng.proj_in.items()}
        self.chead_feat_ok = False
        self.xs = e(nsub + 1, M, H, dtype=f32)
        self.xn = e(nsub, M, H)
        self.rn = e(nsub, M, dtype=f32)
        self.qkv = e(eng.depth, M, 3 * H)
        self.P = e(eng.depth, B * eng.heads, T, self.Tp)
        self.attn = e(eng.depth, M, H)
        self.hpre = e(eng.depth, M, F)
        self.hact = e(eng.depth, M, F)
        self.xnf = e(M, H)
        self.rnf = e(M, dtype=f32)
        # scratch (shared by forward and backward of this workspace)
        self.S = e(B * eng.heads, T, self.Tp, dtype=f32)
        self.dS = e(B * eng.heads, T, self.Tp)
        self.dqkv = e(M, 3 * H)
        self.dh = e(M, F)
        self.dtmp = e(M, H)      # d_attn / d_xn
        self.dx = e(2, M, H, dtype=f32)
        self.dxb = e(2, M, H)
        self.pooled = {}         # t_out -> (xp bf16 (B, t_out, H), d_xp bf16)
        self.in_use = False

    def pooled_bufs(self, eng, t_out):
        if t_out not in self.pooled:
            self.pooled[t_out] = (torch.empty(self.B, t_out, eng.hidden, device=eng.device, dtype=torch.bfloat16),
                                  torch.empty(self.B, t_out, eng.hidden, device=eng.device, dtype=torch.bfloat16))
        return self.pooled[t_out]


class Plan:
    """What one forward call did (needed by its backward)."""

    def __init__(self):
        self.ws = None
        self.dropped = []
        self.present = []       # modalities projected this call
        self.subjects = None    # int64 (B,) on device
        self.pool = True
        self.t_out = None
        self.mode = "predict"   # "predict": readout output; "latents": final-norm output (B, T, H)
        self.x_input = False    # transformer_forward entry: encoder input given as a tensor
        self.training = False
        self.out = None         # optional preallocated (B, O, T') fp32 output (evaluation sweeps write predictions in place)
        self.awaiting = None    # the Engine that counts this plan among its grad-enabled forwards awaiting a backward
        self.keep = True        # a backward may follow: keep what it reads (attention probabilities); False under no_grad

    def settle(self):
        eng, self.awaiting = self.awaiting, None
        if eng is not None:
            eng._grad_plans -= 1

    def __del__(self):
        # a grad-enabled forward whose backward never runs (main.py:349 warm-up call, a skipped loss, an exception):
        # the workspace returns to the pool when autograd drops the graph that holds this plan
        ws = self.__dict__.get("ws")
        if ws is not None:
            ws.in_use = False
        if self.__dict__.get("awaiting") is not None:
            self.settle()


class Engine:
    def __init__(self, model):
        self.model = model
        self.hidden, self.depth, self.heads = model.hidden, model.depth, model.heads
        self.dh = self.hidden // self.heads
        self.ff = 4 * self.hidden
        self.rot = max(self.dh // 2, 32)
        if self.dh % 32 or self.rot % 32 or self.hidden % 64:
            raise TribeError("engine needs head_dim and rotary dim to be multiples of 32 and hidden % 64 == 0")
        self.device = None
        self.flat: FlatParams | None = None
        self._ws_pool = {}
        self._rope = {}
        self.proj_in = {}
        self.comm = None  # set by parallel.GradAllReduce for data-parallel training
        # optimizer-in-backward (optim.TribeAdam.arm_fused_backward): the LAST backward pass of a step applies the Adam
        # update of every GEMM weight inside its wgrad epilogue
        self.fused_opt = None
        self._grad_plans = 0  # grad-enabled forward passes whose backward has not run yet
        # which x_transformers release's arithmetic the encoder follows (oracle/xt_encoder.py): ">= 2.x" = ScaleNorm
        # F.normalize * sqrt(dim) * g and interleaved rotary pairs (fused into the GEMM epilogues); "v1.27" = ScaleNorm
        # x / norm.clamp(1e-5) * g and half-split rotary pairs (a separate in-place pass over q, k / dq, dk)
        self.v127 = getattr(model, "xt_semantics", "v2") == "v1.27"
        self.norm_mult, self.norm_eps = (1.0, 1e-5) if self.v127 else (0.0, 0.0)

    # ------------------------------------------------------------------------------------------------ parameters
    def materialize(self, device):
        """Move every parameter into the flat CUDA buffer (the module owns its device, SURVEY §8b trap 1)."""
        m = self.model
        named = list(m.named_parameters())
        head = [n for n, _ in named if not n.startswith("encoder.")]
        buckets = [head]
        for l in range(self.depth):
            a, f = f"encoder.layers.{2 * l}", f"encoder.layers.{2 * l + 1}"
            buckets.append([f"{a}.1.to_q.weight", f"{a}.1.to_k.weight", f"{a}.1.to_v.weight", f"{a}.1.to_out.weight",
                            f"{f}.1.ff.0.0.weight", f"{f}.1.ff.2.weight", f"{f}.1.ff.0.0.bias", f"{f}.1.ff.2.bias",
                            f"{a}.0.0.g", f"{a}.2.residual_scale", f"{f}.0.0.g", f"{f}.2.residual_scale"])
        buckets[-1].append("encoder.final_norm.g")
        self.device = torch.device(device)
        self.flat = FlatParams(named, buckets, self.device)
        for b in m.buffers():
            b.data = b.data.to(self.device)
        self.proj_in = {mod: lin.in_features for mod, lin in m.projectors.items()}
        self._ws_pool.clear()
        self._rope.clear()  # the rotary table lives on the device too

    def _check_flat(self):
        if self.flat is None or not self.flat.intact():
            if not torch.cuda.is_available():
                raise TribeError("FmriEncoder needs a CUDA device: the B200 kernels have no CPU fallback")
            dev = next((p.device for p in self.model.parameters() if p.is_cuda), torch.device("cuda", torch.cuda.current_device()))
            self.materialize(dev)

    def _rope_table(self, T):
        if T not in self._rope:
            inv = self.model.encoder.rotary_pos_emb.inv_freq.detach().float().cpu()
            ang = torch.arange(T).type_as(inv)[:, None] * inv[None, :]
            self._rope[T] = torch.stack((ang.cos(), ang.sin()), dim=-1).contiguous().to(self.device)
        return self._rope[T]

    MAX_WORKSPACES_PER_SHAPE = 4  # a contrastive train step holds two; anything beyond is a leak, not a need

    def _workspace(self, B, T):
        pool = self._ws_pool.setdefault((B, T), [])
        for ws in pool:
            if not ws.in_use:
                ws.in_use = True
                return ws
        if len(pool) >= self.MAX_WORKSPACES_PER_SHAPE:
            raise TribeError(f"{len(pool)} activation workspaces of shape (B={B}, T={T}) are held by forward passes whose backward never ran "
                             "(outputs with an autograd graph are still referenced); drop them or call under torch.no_grad()")
        ws = Workspace(self, B, T)
        ws.in_use = True
        pool.append(ws)
        return ws

    # ------------------------------------------------------------------------------------------------ helpers
    def _names(self, l):
        a, f = f"encoder.layers.{2 * l}", f"encoder.layers.{2 * l + 1}"
        return a, f

    def _w16(self, name):
        return self.flat.view16(name)

    def _wqkv16(self, a):
        fl, H = self.flat, self.hidden
        o = fl.offsets[f"{a}.1.to_q.weight"]
        return fl.bf16[o: o + 3 * H * H].view(3 * H, H)

    def _p(self, name):
        return self.flat.params[name]

    # ------------------------------------------------------------------------------------------------ forward
    def _validate_batch(self, batch_data):
        """Every modality must be (B, L, D, T) / (B, D, T) with L*D (or D for layer_aggregation="mean") equal to its
        projector's in_features and the same (B, T) as the others; the reference fails in its matmul, the kernels would
        read or write out of bounds instead."""
        m = self.model
        mean = m.config.layer_aggregation == "mean"
        ref = None
        for mod in m.feature_dims:
            if mod not in batch_data or mod not in self.proj_in:
                continue
            x = batch_data[mod]
            if x.dim() not in (3, 4):
                raise TribeError(f"modality '{mod}': expected (B, L, D, T) or (B, D, T), got {tuple(x.shape)}")
            L, D = (1, x.shape[1]) if x.dim() == 3 else (x.shape[1], x.shape[2])
            k = D if mean else L * D
            if k != self.proj_in[mod]:
                raise TribeError(f"modality '{mod}': feature width {k} (L={L}, D={D}) does not match the projector's in_features {self.proj_in[mod]}")
            bt = (x.shape[0], x.shape[-1])
            if ref is None:
                ref = (mod, bt)
            elif bt != ref[1]:
                raise TribeError(f"modality '{mod}' has (B, T) = {bt} but '{ref[0]}' has {ref[1]}")
        if ref is None:
            raise TribeError("batch holds none of the model's modalities")
        return ref[1]

    def forward(self, plan: Plan, batch_data, x_in=None):
        self._check_flat()
        self.flat.refresh_bf16()
        if not plan.x_input:
            self._validate_batch(batch_data)
        try:
            return self._forward(plan, batch_data, x_in)
        except BaseException:
            self.release(plan)
            raise

    def _forward(self, plan: Plan, batch_data, x_in=None):
        m, H, heads, dh, F = self.model, self.hidden, self.heads, self.dh, self.ff
        cfg = m.config
        if plan.x_input:
            B, T, _ = x_in.shape
        else:
            ref = next(batch_data[k] for k in batch_data if k in m.feature_dims)
            B, T = ref.shape[0], ref.shape[-1]
        if T > m.time_pos_embed.shape[1]:
            raise TribeError(f"sequence length {T} exceeds time_pos_embed capacity {m.time_pos_embed.shape[1]}")
        ws = plan.ws = self._workspace(B, T)
        M, Tp = ws.M, ws.Tp
        pos = m.time_pos_embed  # (1, 1024, H) fp32
        x0 = ws.xs[0]

        if plan.x_input:
            xin = x_in.to(self.device, torch.float32).contiguous().view(M, H)
            ops.add_rows_periodic(xin, pos, x0, M, H, T, ld_x=H, ld_pos=H, ld_out=H)
        else:
            mods = list(m.feature_dims.keys())
            cat = cfg.feature_aggregation == "cat"
            width = H // len(mods) if cat else H
            first = True
            plan.present = []
            for i, mod in enumerate(mods):
                col = i * width if cat else 0
                active = mod in m.projectors and mod not in plan.dropped
                if active:
                    ops.ingest_features(batch_data[mod].to(self.device, non_blocking=True), ws.feat[mod], 0, cfg.layer_aggregation == "mean")
                    w16, bias = self._w16(f"projectors.{mod}.weight"), self._p(f"projectors.{mod}.bias")
                    if cat or first:
                        ops.gemm(ops.kmajor(ws.feat[mod]), ops.kmajor(w16), x0, M, width, w16.shape[1], ldd=H, d_off=col, bias=bias,
                                 epilogue=ops.EPI_RESIDUAL, res=pos.view(-1, H)[:, col:], ld_res=H, res_row_mod=T)
                    else:  # "sum": accumulate onto what is already there
                        ops.gemm(ops.kmajor(ws.feat[mod]), ops.kmajor(w16), x0, M, width, w16.shape[1], ldd=H, bias=bias,
                                 epilogue=ops.EPI_RESIDUAL, res=x0, ld_res=H)
                    plan.present.append(mod)
                    first = False
                elif cat:
                    # dropped (model.py:158-159) or projector-less (model.py:143-144) modality: zeros + positional embedding
                    ops.add_rows_periodic(None, pos, x0, M, width, T, ld_pos=H, ld_out=H, pos_off=col, out_off=col)
            if not cat and first:
                ops.add_rows_periodic(None, pos, x0, M, H, T, ld_pos=H, ld_out=H)
        if hasattr(m, "subject_embed"):
            # non-default branch (defaults.py:99 subject_embedding=False): broadcast add of the per-sample embedding row
            emb = m.subject_embed.weight.detach()[plan.subjects]  # (B, H)
            ws.xs[0].view(B, T, H).add_(emb[:, None, :])

        rope = self._rope_table(T)
        scale = dh ** -0.5
        BH = B * heads
        for l in range(self.depth):
            a, f = self._names(l)
            ia, iff = 2 * l, 2 * l + 1
            # ---- attention sub-layer
            ops.scalenorm_fwd(ws.xs[ia], self._p(f"{a}.0.0.g"), ws.xn[ia], ws.rn[ia], self.norm_mult, self.norm_eps)
            qkv = ws.qkv[l]
            if self.v127:
                ops.gemm(ops.kmajor(ws.xn[ia]), ops.kmajor(self._wqkv16(a)), qkv, M, 3 * H, H, ldd=3 * H)
                ops.rope_half(qkv, 0, 2 * heads, dh, self.rot, rope, T)  # q and k heads are adjacent in the packed buffer
            else:
                ops.gemm(ops.kmajor(ws.xn[ia]), ops.kmajor(self._wqkv16(a)), qkv, M, 3 * H, H, ldd=3 * H, epilogue=ops.EPI_ROPE,
                         rope=rope, rope_t=T, rope_dim=self.rot, head_dim=dh, rope_cols=2 * H)
            if ops.attn_fwd_fusable(T, dh):
                # flash-style forward, one launch: S in TMEM -> softmax -> P as the shared-memory A operand of P.V -> O;
                # P goes to HBM only when a backward will read it
                ops.attn_fwd(qkv, 0, H, 2 * H, B, T, heads, dh, scale, ws.attn[l], p_out=ws.P[l] if plan.keep else None)
            else:
                if ops.attn_fusable(T, dh):
                    # P = softmax(q k^T d^-1/2) formed in the tcgen05 epilogue: whole score rows live in TMEM, no fp32 S in HBM
                    ops.attn_scores(qkv, 0, qkv, H, B, T, heads, dh, scale, ws.P[l])
                else:
                    q_op = ops.Operand(qkv, inner=3 * H, rows=T, row_stride=3 * H, batch=B, batch_stride=T * 3 * H, zin_stride=dh, zdiv=heads)
                    k_op = ops.Operand(qkv, inner=3 * H, rows=T, row_stride=3 * H, batch=B, batch_stride=T * 3 * H, inner_off=H, zin_stride=dh,
                                       zdiv=heads)
                    ops.gemm(q_op, k_op, ws.S, T, T, dh, ldd=Tp, batch=BH, z_inner=heads, d_zo=heads * T * Tp, d_zi=T * Tp, alpha=scale)
                    ops.softmax_fwd(ws.S, ws.P[l], T)
                p_op = ops.Operand(ws.P[l], inner=Tp, rows=T, row_stride=Tp, batch=BH, batch_stride=T * Tp)
                v_op = ops.Operand(qkv, inner=3 * H, rows=T, row_stride=3 * H, batch=B, batch_stride=T * 3 * H, mn_major=True,
                                   inner_off=2 * H, zin_stride=dh, zdiv=heads)
                ops.gemm(p_op, v_op, ws.attn[l], T, dh, Tp, ldd=H, batch=BH, z_inner=heads, d_zo=T * H, d_zi=dh)
            ops.gemm(ops.kmajor(ws.attn[l]), ops.kmajor(self._w16(f"{a}.1.to_out.weight")), ws.xs[ia + 1], M, H, H, ldd=H,
                     epilogue=ops.EPI_RESIDUAL, res=ws.xs[ia], ld_res=H, rscale=self._p(f"{a}.2.residual_scale"))
            # ---- feed-forward sub-layer
            ops.scalenorm_fwd(ws.xs[iff], self._p(f"{f}.0.0.g"), ws.xn[iff], ws.rn[iff], self.norm_mult, self.norm_eps)
            ops.gemm(ops.kmajor(ws.xn[iff]), ops.kmajor(self._w16(f"{f}.1.ff.0.0.weight")), ws.hact[l], M, F, H, ldd=F,
                     bias=self._p(f"{f}.1.ff.0.0.bias"), epilogue=ops.EPI_GELU, aux_out=ws.hpre[l] if plan.keep else None, ld_aux=F)  # pre-activation only for a backward
            ops.gemm(ops.kmajor(ws.hact[l]), ops.kmajor(self._w16(f"{f}.1.ff.2.weight")), ws.xs[iff + 1], M, H, F, ldd=H,
                     bias=self._p(f"{f}.1.ff.2.bias"), epilogue=ops.EPI_RESIDUAL, res=ws.xs[iff], ld_res=H,
                     rscale=self._p(f"{f}.2.residual_scale"))
        ops.scalenorm_fwd(ws.xs[2 * self.depth], self._p("encoder.final_norm.g"), ws.xnf, ws.rnf, self.norm_mult, self.norm_eps)

        if plan.mode == "latents":
            lat = torch.empty(B, T, H, device=self.device, dtype=torch.float32)
            ops.cast_bf16_f32(ws.xnf, lat)
            return lat
        return self._readout_fwd(plan, ws)

    def _readout_fwd(self, plan, ws):
        m, H = self.model, self.hidden
        B, T = ws.B, ws.T
        O, S = m.n_outputs, m.predictor.weights.shape[0]
        subj = plan.subjects
        if plan.pool:
            Tq = plan.t_out
            xp, _ = ws.pooled_bufs(self, Tq)
            ops.token_pool_fwd(ws.xnf, xp, B, T, Tq, H)
        else:
            Tq, xp = T, ws.xnf
        out = plan.out
        if out is None:
            out = torch.empty(B, O, Tq, device=self.device, dtype=torch.float32)
        elif tuple(out.shape) != (B, O, Tq) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != self.device:
            raise TribeError(f"forward(out=...): expected contiguous float32 {(B, O, Tq)} on {self.device}, got {tuple(out.shape)} {out.dtype}")
        w16 = self._w16("predictor.weights")
        x_op = ops.Operand(xp, inner=H, rows=Tq, row_stride=H, batch=B, batch_stride=Tq * H)
        w_op = ops.Operand(w16, inner=O, rows=H, row_stride=O, batch=S, batch_stride=H * O, mn_major=True, gather=subj)
        bias = m.predictor.bias
        ops.gemm(x_op, w_op, out, Tq, O, H, ldd=Tq, batch=B, d_zo=O * Tq, transposed=True, bias=bias,
                 bias_gathered=bias is not None, bias_z_stride=O)
        return out

    # ------------------------------------------------------------------------------------------------ backward
    def _grad_target(self, name, acc_flags):
        """fp32 gradient view for ``name`` inside the flat gradient buffer and whether to accumulate into it."""
        fl = self.flat
        p = fl.params[name]
        g = fl.gview(name)
        if p.grad is None:
            acc_flags[name] = False
        elif p.grad.data_ptr() == g.data_ptr():
            acc_flags[name] = True
        else:
            raise TribeError(f"parameter {name} carries a foreign .grad tensor; use zero_grad(set_to_none=True)")
        return g

    def _wgrad(self, a_op, b_op, name, m_, n_, k_, acc, out=None, fused=None, **kw):
        """Weight gradient of ``name`` (or of the adjacent parameters ``name`` lists, written through ``out``).  With
        ``fused`` (the armed optimizer, last backward pass of the step) the Adam update of those parameters happens in the
        GEMM epilogue: the gradient is not stored, ``p.grad`` stays None and ``optimizer.step()`` skips them."""
        names = [name] if isinstance(name, str) else list(name)
        g = out if out is not None else self.flat.gview(names[0])
        adam = fused.fused_wgrad_args(names) if fused is not None else None
        if acc:
            ops.gemm(a_op, b_op, g, m_, n_, k_, ldd=n_, epilogue=ops.EPI_RESIDUAL, res=g, ld_res=n_, res_batched=True, adam=adam, **kw)
        else:
            ops.gemm(a_op, b_op, g, m_, n_, k_, ldd=n_, adam=adam, **kw)
        return adam is not None

    def backward(self, plan: Plan, grad_out: torch.Tensor, want_dx_in=False):
        """Writes parameter gradients straight into the flat gradient buffer (``p.grad`` become views of it; a
        parameter that did not take part — the projector of a dropped modality, model.py:158-159 — keeps ``grad is
        None`` exactly like the reference's autograd) and returns d(loss)/d(encoder input) when asked."""
        fl, m, ws = self.flat, self.model, plan.ws
        if ws is None:
            raise TribeError("backward through this forward pass ran already: its activations were released "
                             "(backward(retain_graph=True) followed by a second backward is not supported)")
        H, heads, dh, F = self.hidden, self.heads, self.dh, self.ff
        B, T, M, Tp = ws.B, ws.T, ws.M, ws.Tp
        BH = B * heads
        fl.ensure_grad()
        touched = []
        acc = {}
        last_pass = self._grad_plans <= 1  # no other grad-enabled forward is still waiting for its backward
        plan.settle()
        fo = self.fused_opt if (last_pass and self.comm is None and plan.training) else None

        def tgt(name):
            g = self._grad_target(name, acc)
            touched.append(name)
            return g

        def wgrad(a_op, b_op, name, m_, n_, k_, **kw):
            names = [name] if isinstance(name, str) else list(name)
            for n in names:
                tgt(n)
            if self._wgrad(a_op, b_op, names, m_, n_, k_, acc[names[0]], fused=fo, **kw):
                for n in names:  # updated in the epilogue: no gradient to publish (optimizer.step() skips grad-less parameters)
                    touched.remove(n)
                    fl.params[n].grad = None

        def publish():
            # hand the finished gradients to autograd's view of the parameters (p.grad = view of the flat buffer);
            # done bucket by bucket so that an overlapped all-reduce / optimizer step can consume a finished layer
            # while the backward of earlier layers is still running
            while touched:
                n = touched.pop()
                p = fl.params[n]
                if p.grad is None:
                    p.grad = fl.gview(n)

        dx_cur, dx_nxt = ws.dx[0], ws.dx[1]
        dxb_cur, dxb_nxt = ws.dxb[0], ws.dxb[1]
        gfin = tgt("encoder.final_norm.g")
        if not acc["encoder.final_norm.g"]:
            ops.zero_(gfin)

        if plan.mode == "latents":
            d_xnf = ws.dtmp
            ops.cast_f32_bf16(grad_out.contiguous().float().view(-1), d_xnf.view(-1))
        else:
            O, S = m.n_outputs, m.predictor.weights.shape[0]
            Tq = plan.t_out if plan.pool else T
            go = grad_out.contiguous().float()
            dyT = torch.empty(B, Tq, O, device=self.device, dtype=torch.bfloat16)
            ops.transpose_cast_bot(go, dyT)
            xp, d_xp = ws.pooled_bufs(self, Tq) if plan.pool else (ws.xnf, ws.dtmp)
            w16 = self._w16("predictor.weights")
            # dgrad: d_xp[b] (Tq, H) = dy[b] (Tq, O) @ W[s_b]^T  (W[s] (H, O): K-major B with K = O)
            dy_op = ops.Operand(dyT, inner=O, rows=Tq, row_stride=O, batch=B, batch_stride=Tq * O)
            w_op = ops.Operand(w16, inner=O, rows=H, row_stride=O, batch=S, batch_stride=H * O, gather=plan.subjects)
            ops.gemm(dy_op, w_op, d_xp, Tq, H, O, ldd=H, batch=B, d_zo=Tq * H)
            # grouped wgrad: dW[s] (H, O) = sum_{b: s_b = s} xp[b]^T dy[b]
            xa = ops.Operand(xp, inner=H, rows=Tq, row_stride=H, batch=B, batch_stride=Tq * H, mn_major=True)
            dyb = ops.Operand(dyT, inner=O, rows=Tq, row_stride=O, batch=B, batch_stride=Tq * O, mn_major=True)
            wgrad(xa, dyb, "predictor.weights", H, O, Tq, out=fl.gview("predictor.weights").view(S * H, O), batch=S, d_zo=H * O,
                  kgroup=plan.subjects)
            if m.predictor.bias is not None:
                gB = tgt("predictor.bias")
                if not acc["predictor.bias"]:
                    ops.zero_(gB)
                ops.subject_bias_grad(dyT, plan.subjects, gB, B, Tq, O, S)
            if plan.pool:
                d_xnf = ws.dtmp
                ops.token_pool_bwd(d_xp, d_xnf, B, T, Tq, H)
            else:
                d_xnf = d_xp
        # final ScaleNorm backward
        ops.sublayer_bwd(None, d_xnf, ws.xs[2 * self.depth], ws.rnf, self._p("encoder.final_norm.g"), None, dx_cur, dxb_cur, None, gfin,
                         self.norm_mult)

        rope = self._rope_table(T)
        scale = dh ** -0.5
        for l in reversed(range(self.depth)):
            a, f = self._names(l)
            ia, iff = 2 * l, 2 * l + 1
            small = [f"{f}.1.ff.0.0.bias", f"{f}.1.ff.2.bias", f"{a}.0.0.g", f"{a}.2.residual_scale", f"{f}.0.0.g", f"{f}.2.residual_scale"]
            gs = {n: tgt(n) for n in small}
            atomics = (f"{a}.0.0.g", f"{a}.2.residual_scale", f"{f}.0.0.g", f"{f}.2.residual_scale")  # adjacent in the flat layout
            if not any(acc[n] for n in atomics):
                lo, hi = fl.offsets[atomics[0]], fl.offsets[atomics[-1]] + H
                ops.zero_(fl.grad[lo:hi])  # one fill per layer for every atomically accumulated small gradient
            else:
                for n in atomics:
                    if not acc[n]:
                        ops.zero_(gs[n])
            # ================= feed-forward sub-layer (x_out = FF(norm(x_in)) + x_in * rs)
            w2, w1 = self._w16(f"{f}.1.ff.2.weight"), self._w16(f"{f}.1.ff.0.0.weight")
            # d_hpre = (dx @ W2) * gelu'(hpre)          W2 (H, F): MN-major B (N = F contiguous, K = H rows)
            ops.gemm(ops.kmajor(dxb_cur), ops.mnmajor(w2), ws.dh, M, F, H, ldd=F, epilogue=ops.EPI_GELU_BWD, aux_in=ws.hpre[l], ld_aux=F)
            wgrad(ops.mnmajor(dxb_cur), ops.mnmajor(ws.hact[l]), f"{f}.1.ff.2.weight", H, F, M)
            ops.colsum(dx_cur, gs[f"{f}.1.ff.2.bias"], accumulate=acc[f"{f}.1.ff.2.bias"])
            # d_xn = d_hpre @ W1                         W1 (F, H): MN-major B (N = H contiguous, K = F rows)
            ops.gemm(ops.kmajor(ws.dh), ops.mnmajor(w1), ws.dtmp, M, H, F, ldd=H)
            wgrad(ops.mnmajor(ws.dh), ops.mnmajor(ws.xn[iff]), f"{f}.1.ff.0.0.weight", F, H, M)
            ops.colsum(ws.dh, gs[f"{f}.1.ff.0.0.bias"], accumulate=acc[f"{f}.1.ff.0.0.bias"])
            ops.sublayer_bwd(dx_cur, ws.dtmp, ws.xs[iff], ws.rn[iff], self._p(f"{f}.0.0.g"), self._p(f"{f}.2.residual_scale"),
                             dx_nxt, dxb_nxt, gs[f"{f}.2.residual_scale"], gs[f"{f}.0.0.g"], self.norm_mult)
            dx_cur, dx_nxt, dxb_cur, dxb_nxt = dx_nxt, dx_cur, dxb_nxt, dxb_cur
            # ================= attention sub-layer
            wo = self._w16(f"{a}.1.to_out.weight")
            d_attn = ws.dtmp
            ops.gemm(ops.kmajor(dxb_cur), ops.mnmajor(wo), d_attn, M, H, H, ldd=H)
            wgrad(ops.mnmajor(dxb_cur), ops.mnmajor(ws.attn[l]), f"{a}.1.to_out.weight", H, H, M)
            qkv = ws.qkv[l]
            # dP = dO V^T  (per (b, h); both K-major over head dims)
            if ops.attn_fusable(T, dh):
                # dS = P o (dP - rowsum(dP o P)) d^-1/2 with dP = dO V^T never leaving TMEM
                ops.attn_scores(d_attn, 0, qkv, 2 * H, B, T, heads, dh, scale, ws.dS, p_in=ws.P[l])
            else:
                do_op = ops.Operand(d_attn, inner=H, rows=T, row_stride=H, batch=B, batch_stride=T * H, zin_stride=dh, zdiv=heads)
                v_op = ops.Operand(qkv, inner=3 * H, rows=T, row_stride=3 * H, batch=B, batch_stride=T * 3 * H, inner_off=2 * H, zin_stride=dh,
                                   zdiv=heads)
                ops.gemm(do_op, v_op, ws.S, T, T, dh, ldd=Tp, batch=BH, z_inner=heads, d_zo=heads * T * Tp, d_zi=T * Tp)
                ops.softmax_bwd(ws.P[l], ws.S, ws.dS, scale, T)
            # dV = P^T dO
            pt_op = ops.Operand(ws.P[l], inner=Tp, rows=T, row_stride=Tp, batch=BH, batch_stride=T * Tp, mn_major=True)
            dom_op = ops.Operand(d_attn, inner=H, rows=T, row_stride=H, batch=B, batch_stride=T * H, mn_major=True, zin_stride=dh, zdiv=heads)
            ops.gemm(pt_op, dom_op, ws.dqkv, T, dh, T, ldd=3 * H, batch=BH, z_inner=heads, d_zo=T * 3 * H, d_zi=dh, d_off=2 * H)
            # dQ = dS K (inverse rotary in the epilogue), dK = dS^T Q
            ds_op = ops.Operand(ws.dS, inner=Tp, rows=T, row_stride=Tp, batch=BH, batch_stride=T * Tp)
            km_op = ops.Operand(qkv, inner=3 * H, rows=T, row_stride=3 * H, batch=B, batch_stride=T * 3 * H, mn_major=True, inner_off=H,
                                zin_stride=dh, zdiv=heads)
            unrope = {} if self.v127 else dict(epilogue=ops.EPI_ROPE, rope=rope, rope_t=T, rope_dim=self.rot, head_dim=dh, rope_cols=dh, rope_sign=-1.0)
            ops.gemm(ds_op, km_op, ws.dqkv, T, dh, Tp, ldd=3 * H, batch=BH, z_inner=heads, d_zo=T * 3 * H, d_zi=dh, **unrope)
            dst_op = ops.Operand(ws.dS, inner=Tp, rows=T, row_stride=Tp, batch=BH, batch_stride=T * Tp, mn_major=True)
            qm_op = ops.Operand(qkv, inner=3 * H, rows=T, row_stride=3 * H, batch=B, batch_stride=T * 3 * H, mn_major=True, zin_stride=dh, zdiv=heads)
            ops.gemm(dst_op, qm_op, ws.dqkv, T, dh, T, ldd=3 * H, batch=BH, z_inner=heads, d_zo=T * 3 * H, d_zi=dh, d_off=H, **unrope)
            if self.v127:
                ops.rope_half(ws.dqkv, 0, 2 * heads, dh, self.rot, rope, T, sign=-1.0)  # transpose of the rotation on dq and dk
            # d_xn = dqkv @ Wqkv ;  dWqkv = dqkv^T xn
            wqkv = self._wqkv16(a)
            ops.gemm(ops.kmajor(ws.dqkv), ops.mnmajor(wqkv), ws.dtmp, M, H, 3 * H, ldd=H)
            o = fl.offsets[f"{a}.1.to_q.weight"]
            gqkv = fl.grad[o: o + 3 * H * H].view(3 * H, H)
            wgrad(ops.mnmajor(ws.dqkv), ops.mnmajor(ws.xn[ia]), [f"{a}.1.{n}.weight" for n in ("to_q", "to_k", "to_v")], 3 * H, H, M, out=gqkv)
            ops.sublayer_bwd(dx_cur, ws.dtmp, ws.xs[ia], ws.rn[ia], self._p(f"{a}.0.0.g"), self._p(f"{a}.2.residual_scale"),
                             dx_nxt, dxb_nxt, gs[f"{a}.2.residual_scale"], gs[f"{a}.0.0.g"], self.norm_mult)
            dx_cur, dx_nxt, dxb_cur, dxb_nxt = dx_nxt, dx_cur, dxb_nxt, dxb_cur
            if self.comm is not None:
                publish()
                self.comm.bucket_ready(l + 1)

        # ---- encoder input: positional embedding, projectors
        gpos = tgt("time_pos_embed")
        if not acc["time_pos_embed"]:
            ops.zero_(gpos)
        ops.colsum(dx_cur.view(B, T * H), gpos.view(-1)[: T * H], accumulate=acc["time_pos_embed"])
        if hasattr(m, "subject_embed"):
            gE = tgt("subject_embed.weight")
            if not acc["subject_embed.weight"]:
                ops.zero_(gE)
            gE.index_add_(0, plan.subjects, dx_cur.view(B, T, H).sum(1))
        if not plan.x_input:
            mods = list(m.feature_dims.keys())
            cat = m.config.feature_aggregation == "cat"
            width = H // len(mods) if cat else H
            for i, mod in enumerate(mods):
                if mod not in plan.present:
                    continue
                col = i * width if cat else 0
                wn, bn = f"projectors.{mod}.weight", f"projectors.{mod}.bias"
                gb = tgt(bn)
                K = self.proj_in[mod]
                a_op = ops.Operand(dxb_cur, inner=H, rows=M, row_stride=H, mn_major=True, inner_off=col)
                wgrad(a_op, ops.mnmajor(ws.feat[mod]), wn, width, K, M)
                ops.colsum(dx_cur[:, col: col + width], gb, accumulate=acc[bn])
        publish()
        if self.comm is not None:
            self.comm.bucket_ready(0)
        dx_ret = dx_cur.view(B, T, H).clone() if want_dx_in else None
        self.release(plan)
        return dx_ret

    def release(self, plan):
        if plan.ws is not None:
            plan.ws.in_use = False
            plan.ws = None
