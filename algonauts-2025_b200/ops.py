"""Thin torch-tensor wrappers over the C ABI (``include/tribe_b200.h``).  Every function launches hand-written sm_100a
kernels on ``torch.cuda.current_stream()``; tensors must live on the current CUDA device.  No CPU fallback."""
from __future__ import annotations

import ctypes
import dataclasses

import torch

from . import _lib
from ._lib import TribeError, TribeGemm, TribeOperand, check

EPI_STORE, EPI_GELU, EPI_RESIDUAL, EPI_GELU_BWD, EPI_ROPE = 0, 1, 2, 3, 4
_DT = {torch.float32: 0, torch.float64: 1, torch.bfloat16: 2, torch.float16: 3}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise TribeError("tribe ops need CUDA tensors (there is no CPU path)")
    return ctypes.c_void_p(t.data_ptr())


_DEBUG_CAPTURE = bool(int(__import__("os").environ.get("TRIBE_DEBUG_CAPTURE", "0")))


_DEBUG_LASTERR = bool(int(__import__("os").environ.get("TRIBE_DEBUG_LASTERR", "0")))


def peek_last_error(where: str) -> None:
    """Debugging aid: raise when the CUDA runtime's (non-sticky) last-error slot is set — some earlier call failed without
    anyone consuming its error, and torch would report it at an unrelated later call."""
    rc = int(_lib.load().tribe_peek_last_error())
    if rc:
        raise TribeError(f"CUDA last-error {rc} is set {where}")


def drain_stale_error(where: str) -> int:
    """Clear a non-sticky CUDA error somebody else left in the runtime's last-error slot (every call of this library
    consumes its own), so that torch does not report it at the synchronisation that follows; warns with the code."""
    rc = int(_lib.load().tribe_take_last_error())
    if rc:
        import warnings

        warnings.warn(f"cleared a stale CUDA runtime error ({rc}) found {where}; it was not raised by a tribe kernel", RuntimeWarning, stacklevel=2)
    return rc


def _run(name, *args):
    if _DEBUG_LASTERR:
        peek_last_error(f"BEFORE {name} (left behind by torch or by an earlier call)")
    rc = getattr(_lib.load(), name)(*args)
    if rc:
        check(rc, name)
    if _DEBUG_LASTERR:
        peek_last_error(f"AFTER {name}")
    if _DEBUG_CAPTURE:  # bisecting a broken CUDA-graph capture: the stream status turns into an error right after the culprit
        try:
            torch.cuda.is_current_stream_capturing()
        except Exception as e:  # noqa: BLE001
            raise TribeError(f"capture invalid after {name}: {e}") from e


def _need(t, dtype, name):
    if t.dtype != dtype or not t.is_cuda or not t.is_contiguous():
        raise TribeError(f"{name}: expected contiguous CUDA {dtype}, got {t.dtype} {t.device} contiguous={t.is_contiguous()}")


@dataclasses.dataclass
class Operand:
    """bf16 GEMM operand viewed as (inner, rows, batch); see TribeOperand in include/tribe_b200.h."""

    t: torch.Tensor
    inner: int
    rows: int
    row_stride: int
    batch: int = 1
    batch_stride: int = 0
    mn_major: bool = False
    inner_off: int = 0
    zin_stride: int = 0
    zdiv: int = 1
    gather: torch.Tensor | None = None
    elem_off: int = 0  # element offset added to the base pointer

    def c(self) -> TribeOperand:
        if self.t.dtype != torch.bfloat16:
            raise TribeError("GEMM operands must be bf16")
        o = TribeOperand()
        o.ptr = self.t.data_ptr() + 2 * self.elem_off
        o.inner, o.rows, o.batch = self.inner, self.rows, self.batch
        o.row_stride, o.batch_stride = self.row_stride, self.batch_stride
        o.mn_major, o.inner_off, o.zin_stride, o.zdiv = int(self.mn_major), self.inner_off, self.zin_stride, self.zdiv
        o.gather = self.gather.data_ptr() if self.gather is not None else None
        return o


def kmajor(t: torch.Tensor, **kw) -> Operand:
    """2-D row-major (rows, K) bf16 tensor as a K-major operand."""
    assert t.dim() == 2 and t.stride(1) == 1
    return Operand(t, inner=t.shape[1], rows=t.shape[0], row_stride=t.stride(0), **kw)


def mnmajor(t: torch.Tensor, **kw) -> Operand:
    """2-D row-major (K, MN) bf16 tensor as an MN-major operand (contraction along rows)."""
    assert t.dim() == 2 and t.stride(1) == 1
    return Operand(t, inner=t.shape[1], rows=t.shape[0], row_stride=t.stride(0), mn_major=True, **kw)


def gemm(a: Operand, b: Operand, out: torch.Tensor, m: int, n: int, k: int, *, ldd: int, batch: int = 1, z_inner: int = 1,
         d_zo: int = 0, d_zi: int = 0, d_off: int = 0, transposed: bool = False, epilogue: int = EPI_STORE, alpha: float = 1.0,
         bias: torch.Tensor | None = None, bias_gathered: bool = False, bias_z_stride: int = 0,
         res: torch.Tensor | None = None, ld_res: int = 0, res_row_mod: int = 0, res_batched: bool = False, rscale: torch.Tensor | None = None,
         aux_in: torch.Tensor | None = None, aux_out: torch.Tensor | None = None, ld_aux: int = 0,
         rope: torch.Tensor | None = None, rope_t: int = 0, rope_dim: int = 0, head_dim: int = 0, rope_cols: int = 0,
         rope_sign: float = 1.0, kgroup: torch.Tensor | None = None, block_n: int = 0, probe=None, adam=None) -> None:
    """D[z] = epilogue(alpha * A[z] @ B[z]^T) on the tcgen05 GEMM.  ``out`` is bf16 or fp32.
    ``adam``: optional ``(p_ptr, m_ptr, v_ptr, shadow_ptr, hyper_ptr, keep_grad)`` — device addresses of the fp32 master
    weights / Adam moments / bf16 shadow laid out like ``out`` (offset ``d_off`` included by the caller) and of the device
    hyper-parameter block: the epilogue applies the optimizer step to the finished gradient tile (wgrad GEMMs)."""
    if out.dtype not in (torch.bfloat16, torch.float32) or not out.is_cuda:
        raise TribeError("gemm output must be a CUDA bf16/fp32 tensor")
    g = TribeGemm()
    g.a, g.b = a.c(), b.c()
    g.m, g.n, g.k, g.batch, g.z_inner = m, n, k, batch, z_inner
    if kgroup is not None:
        g.kgroup, g.kgroup_len = kgroup.data_ptr(), kgroup.numel()
    g.d = out.data_ptr() + d_off * out.element_size()
    g.d_f32, g.d_transposed = int(out.dtype == torch.float32), int(transposed)
    g.ldd, g.d_zo_stride, g.d_zi_stride = ldd, d_zo, d_zi
    g.epilogue, g.alpha = epilogue, alpha
    for name, t, dt in (("bias", bias, torch.float32), ("res", res, torch.float32), ("rscale", rscale, torch.float32),
                        ("aux_in", aux_in, torch.bfloat16), ("aux_out", aux_out, torch.bfloat16), ("rope", rope, torch.float32)):
        if t is not None:
            if t.dtype != dt or not t.is_cuda:
                raise TribeError(f"gemm {name}: expected CUDA {dt}")
            setattr(g, name, t.data_ptr())
    g.bias_gathered, g.bias_z_stride = int(bias_gathered), bias_z_stride
    g.ld_res, g.res_row_mod, g.res_batched, g.ld_aux = ld_res, res_row_mod, int(res_batched), ld_aux
    g.rope_t, g.rope_dim, g.head_dim, g.rope_cols, g.rope_sign = rope_t, rope_dim, head_dim, rope_cols, rope_sign
    g.block_n = block_n
    if adam is not None:
        g.adam_p, g.adam_m, g.adam_v, g.adam_shadow, g.adam_hyper, g.adam_keep_grad = (int(adam[0]), int(adam[1]), int(adam[2]), int(adam[3]) or None,
                                                                                       int(adam[4]), int(bool(adam[5])))
    if SPLITK:
        ws = _splitk_workspace(out.device)
        g.splitk_ws, g.splitk_ws_bytes = ws.data_ptr(), ws.numel()
    log = GEMM_LOG
    if log is not None:
        # inside a CUDA-graph capture the pair becomes event-record NODES (external events): every replay re-records
        # them, so per-launch durations can be read after a replay of the whole step
        ext = torch.cuda.is_current_stream_capturing()
        e0, e1 = torch.cuda.Event(enable_timing=True, external=ext), torch.cuda.Event(enable_timing=True, external=ext)
        e0.record()
    if probe is None:
        _run("tribe_gemm_bf16", ctypes.byref(g), _stream())
    else:
        _run("tribe_gemm_bf16_probe", ctypes.byref(g), _stream(), *probe)
    if log is not None:
        e1.record()
        log.append((e0, e1, 2.0 * m * n * k * (kgroup.numel() if kgroup is not None else batch)))


# bench.py sets this to a list to collect (start event, end event, algorithmic FLOPs) per GEMM launch
GEMM_LOG = None

# Split-K tail scheduling (ragged last wave of tiles split along K over idle SMs); one zeroed 20 MiB workspace per
# device, used by the GEMMs of the current stream only (the kernel leaves it zeroed).
SPLITK = True
_SPLITK_WS: dict = {}


def _splitk_workspace(device) -> torch.Tensor:
    key = (device.type, device.index)
    ws = _SPLITK_WS.get(key)
    if ws is None:
        ws = _SPLITK_WS[key] = torch.zeros(20 * 1024 * 1024 + 1024, device=device, dtype=torch.uint8)
    return ws


def linear(x: torch.Tensor, w: torch.Tensor, out: torch.Tensor, **kw) -> None:
    """out (M, N) = x (M, K) @ w (N, K)^T (+ epilogue): the nn.Linear forward shape."""
    gemm(kmajor(x), kmajor(w), out, x.shape[0], w.shape[0], x.shape[1], ldd=kw.pop("ldd", out.stride(0)), **kw)


def ingest_features(x: torch.Tensor, out: torch.Tensor, col_off: int, layer_mean: bool) -> None:
    """(B, L, D, T) any float dtype -> bf16 rows of out (B*T, ld) at column col_off (model.py:147-155)."""
    if x.dim() == 3:
        x = x.unsqueeze(1)
    x = x.contiguous()
    B, L, D, T = x.shape
    _need(out, torch.bfloat16, "ingest out")
    _run("tribe_ingest_features", _ptr(x), _DT[x.dtype], B, L, D, T, int(layer_mean), _ptr(out), out.stride(0), col_off,
                                            _stream())


def scalenorm_fwd(x, g, y, rnorm, gain_mult: float = 0.0, eps: float = 0.0) -> None:
    """y = x / max(||x||, eps) * gain_mult * g; the zeros select sqrt(dim) and 1e-12 (x_transformers >= 2.x ScaleNorm)."""
    _need(x, torch.float32, "scalenorm x"), _need(y, torch.bfloat16, "scalenorm y")
    rows, dim = x.shape
    _run("tribe_scalenorm_fwd", _ptr(x), _ptr(g), _ptr(y), _ptr(rnorm), rows, dim, float(gain_mult), float(eps), _stream())


def sublayer_bwd(dy_out, d_xn, x_in, rnorm, g, rs, dx_in, dx_in_bf16, d_rs, d_g, gain_mult: float = 0.0) -> None:
    rows, dim = x_in.shape
    _run("tribe_sublayer_bwd", _ptr(dy_out), _ptr(d_xn), _ptr(x_in), _ptr(rnorm), _ptr(g), _ptr(rs), _ptr(dx_in),
                                         _ptr(dx_in_bf16), _ptr(d_rs), _ptr(d_g), rows, dim, float(gain_mult), _stream())


def rope_half(x, col_off: int, n_heads: int, head_dim: int, rot_dim: int, table, T: int, sign: float = 1.0) -> None:
    """Half-split rotary (x_transformers 1.27.x) in place on the heads at columns [col_off, col_off + n_heads*head_dim) of
    the bf16 (rows, ld) buffer ``x``; ``table`` is the engine's (T, rot_dim/2, 2) cos/sin table; sign=-1: transpose."""
    _need(x, torch.bfloat16, "rope_half x"), _need(table, torch.float32, "rope_half table")
    if x.dim() != 2 or x.stride(1) != 1:
        raise TribeError("rope_half: 2-D row-major buffer expected")
    _run("tribe_rope_half", _ptr(x), x.shape[0], x.stride(0), col_off, n_heads, head_dim, rot_dim, _ptr(table), T, float(sign), _stream())


def softmax_fwd(s, p, n_valid) -> None:
    _need(s, torch.float32, "softmax s"), _need(p, torch.bfloat16, "softmax p")
    ld = s.shape[-1]
    _run("tribe_softmax_fwd", _ptr(s), _ptr(p), s.numel() // ld, n_valid, ld, _stream())


def softmax_bwd(p, dp, ds, scale, n_valid) -> None:
    ld = p.shape[-1]
    _run("tribe_softmax_bwd", _ptr(p), _ptr(dp), _ptr(ds), scale, p.numel() // ld, n_valid, ld, _stream())


def colsum(x, out, y=None, accumulate=False) -> None:
    """out[c] (+)= sum_r x[r, c] (* y[r, c])."""
    rows, cols = x.shape
    _run("tribe_colsum", _ptr(x), _DT[x.dtype], _ptr(y), _DT[y.dtype] if y is not None else 0, _ptr(out), rows, cols,
                                   x.stride(0), int(accumulate), _stream())


def cast_f32_bf16(src, dst) -> None:
    _need(src, torch.float32, "cast src"), _need(dst, torch.bfloat16, "cast dst")
    _run("tribe_cast_f32_bf16", _ptr(src), _ptr(dst), src.numel(), _stream())


def axpby(src, dst, a=1.0, accumulate=False) -> None:
    _run("tribe_axpby_f32", _ptr(src), _ptr(dst), a, int(accumulate), src.numel(), _stream())


def adaptive_avg_pool_fwd(x, t_out) -> torch.Tensor:
    """fp32 (..., T) -> (..., t_out) with nn.AdaptiveAvgPool1d windows (model.py:60)."""
    x = x.contiguous()
    _need(x, torch.float32, "pool x")
    y = torch.empty(*x.shape[:-1], t_out, device=x.device, dtype=torch.float32)
    _run("tribe_adaptive_avg_pool_fwd", _ptr(x), _ptr(y), x.numel() // x.shape[-1], x.shape[-1], t_out, _stream())
    return y


def adaptive_avg_pool_bwd(dy, t_in) -> torch.Tensor:
    dy = dy.contiguous()
    _need(dy, torch.float32, "pool dy")
    dx = torch.empty(*dy.shape[:-1], t_in, device=dy.device, dtype=torch.float32)
    _run("tribe_adaptive_avg_pool_bwd", _ptr(dy), _ptr(dx), dy.numel() // dy.shape[-1], t_in, dy.shape[-1], _stream())
    return dx


def token_pool_fwd(x, y, B, t_in, t_out, C) -> None:
    _run("tribe_token_pool_fwd", _ptr(x), _ptr(y), B, t_in, t_out, C, _stream())


def token_pool_bwd(dy, dx, B, t_in, t_out, C) -> None:
    _run("tribe_token_pool_bwd", _ptr(dy), _DT[dy.dtype], _ptr(dx), _DT[dx.dtype], B, t_in, t_out, C, _stream())


def add_rows_periodic(x, pos, out, rows, cols, row_mod, *, ld_x=0, ld_pos, ld_out, x_off=0, pos_off=0, out_off=0) -> None:
    """out[r, c] = (x[r, c] if x is not None else 0) + pos[r % row_mod, c]; offsets are in elements."""
    xp = ctypes.c_void_p(x.data_ptr() + 4 * x_off) if x is not None else None
    _run("tribe_add_rows_periodic", xp, ld_x, ctypes.c_void_p(pos.data_ptr() + 4 * pos_off), ld_pos,
                                              ctypes.c_void_p(out.data_ptr() + 4 * out_off), ld_out, rows, cols, row_mod, _stream())


def transpose_cast_bot(x, y) -> None:
    B, O, T = x.shape
    _need(x, torch.float32, "transpose x")
    _run("tribe_transpose_cast_bot", _ptr(x), _ptr(y), B, O, T, _stream())


def subject_bias_grad(dy, subjects, d_bias, B, T, O, n_subjects) -> None:
    _run("tribe_subject_bias_grad", _ptr(dy), _ptr(subjects), _ptr(d_bias), B, T, O, n_subjects, _stream())


def check_subjects(subjects, n_subjects, flag, clamped=None) -> None:
    _run("tribe_check_subjects", _ptr(subjects), subjects.numel(), n_subjects, _ptr(flag), _ptr(clamped), _stream())


def mse_fwd_bwd(pred, target, want_grad=True, grad_scale=1.0):
    """nn.MSELoss forward (+ gradient wrt pred in the same pass).  Returns (loss[1] fp32, grad or None)."""
    _need(pred, torch.float32, "mse pred"), _need(target, torch.float32, "mse target")
    loss = torch.empty(1, device=pred.device, dtype=torch.float32)
    grad = torch.empty_like(pred) if want_grad else None
    partial = torch.empty(1024, device=pred.device, dtype=torch.float64)
    _run("tribe_mse_fwd_bwd", _ptr(pred), _ptr(target), _ptr(loss), _ptr(grad), grad_scale, pred.numel(), _ptr(partial),
                                        _stream())
    return loss, grad


def _bdt_view(pred, target):
    """(B, D, T) operands as the stats kernels address them; a parcel slice of a contiguous tensor is read in place."""
    b, d, t = pred.shape
    for x, name in ((pred, "pearson pred"), (target, "pearson target")):
        if x.dtype != torch.float32 or not x.is_cuda or x.shape != pred.shape:
            raise TribeError(f"{name}: expected CUDA float32 of shape {tuple(pred.shape)}, got {x.dtype} {x.device} {tuple(x.shape)}")
    if not (pred.stride(2) == 1 and pred.stride(1) == t and pred.stride() == target.stride()) and b * d * t > 0:
        pred, target = pred.contiguous(), target.contiguous()
    return pred, target


def pearson_pick_shift(pred, target, shift, *, layout: str) -> None:
    """shift fp32 (2, O) <- the first sample of every parcel of pred / target (per-parcel pivots for ``pearson_stats``)."""
    _need(shift, torch.float32, "pearson shift")
    if layout == "no":
        _need(pred, torch.float32, "pearson pred"), _need(target, torch.float32, "pearson target")
        o, stride_p = pred.shape[1], 1
    else:
        pred, target = _bdt_view(pred, target)
        o, stride_p = pred.shape[1], pred.shape[2]
    if shift.shape != (2, o):
        raise TribeError(f"pearson shift must be (2, {o})")
    if pred.numel():
        _run("tribe_pearson_pick_shift", _ptr(pred), _ptr(target), o, stride_p, _ptr(shift), _stream())


def pearson_recenter(stats, shift_old, shift_new) -> None:
    """In place: statistics about pivots ``shift_old`` -> about ``shift_new`` (None = 0); stats fp64 (G, 6, O)."""
    g, _, o = stats.shape
    _run("tribe_pearson_recenter", _ptr(stats), g, o, _ptr(shift_old), _ptr(shift_new), _stream())


def pearson_stats(pred, target, stats, *, layout: str, group=None, n_groups: int = 1, shift=None) -> None:
    """Accumulate per-parcel sufficient statistics into ``stats`` (fp64 [n_groups, 6, O]); ``shift`` fp32 (2, O): per-parcel
    pivots subtracted before the products (every call accumulating into one block must pass the same ones).

    layout "no": pred/target are row-major (N, O).  layout "bdt": they are (B, D, T) and rows are the flattened
    ``(b t)`` of ``rearrange(x, "b d t -> (b t) d")`` (pl_module.py:54-55, main.py:472-473) without materialising it."""
    if stats.dtype != torch.float64:
        raise TribeError("pearson stats must be fp64")
    if layout == "no":
        _need(pred, torch.float32, "pearson pred"), _need(target, torch.float32, "pearson target")
        n, o = pred.shape
        args = (n, o, 1, o, 1, 1)
    elif layout == "bdt":
        # a parcel slice x[:, lo:hi] of a contiguous (B, D, T) tensor is read in place (stride_b stays D_full * T)
        b, d, t = pred.shape
        pred, target = _bdt_view(pred, target)
        args = (b * t, d, t, pred.stride(0) if b > 1 else d * t, t, 1)
    else:
        raise TribeError(f"unknown layout {layout}")
    if shift is not None and (shift.dtype != torch.float32 or tuple(shift.shape) != (2, args[1]) or not shift.is_contiguous()):
        raise TribeError(f"pearson shift must be contiguous float32 (2, {args[1]})")
    _run("tribe_pearson_stats", _ptr(pred), _ptr(target), *args, _ptr(group), n_groups if group is not None else 0,
                                          _ptr(shift), _ptr(stats), _stream())


def pearson_r(pred, target, *, layout: str, want_mean=False):
    """One-shot per-parcel Pearson r of a whole (N, O) / (B, O, T) pair: pivots from the first row, pivoted statistics,
    finalize -> (r (O,), mean or None)."""
    o = pred.shape[1]
    shift = torch.empty(2, o, device=pred.device, dtype=torch.float32)
    stats = torch.zeros(1, 6, o, device=pred.device, dtype=torch.float64)
    pearson_pick_shift(pred, target, shift, layout=layout)
    pearson_stats(pred, target, stats, layout=layout, shift=shift)
    return pearson_finalize(stats[0], want_mean=want_mean)


def pearson_finalize(stats_one_group, want_mean=False):
    o = stats_one_group.shape[-1]
    r = torch.empty(o, device=stats_one_group.device, dtype=torch.float32)
    mean = torch.empty(1, device=r.device, dtype=torch.float32) if want_mean else None
    _run("tribe_pearson_finalize", _ptr(stats_one_group), o, _ptr(r), _ptr(mean), _stream())
    return r, mean


def cast_bf16_f32(src, dst) -> None:
    _run("tribe_cast_bf16_f32", _ptr(src), _ptr(dst), src.numel(), _stream())


def nce_expsums(logits, n, shift, row_sum, col_sum) -> None:
    _run("tribe_nce_expsums", _ptr(logits), n, logits.stride(0), shift, _ptr(row_sum), _ptr(col_sum), _stream())


def nce_loss(logits, n, shift, row_sum, col_sum, loss) -> None:
    _run("tribe_nce_loss", _ptr(logits), n, logits.stride(0), shift, _ptr(row_sum), _ptr(col_sum), _ptr(loss), _stream())


def nce_grad(logits, n, shift, row_sum, col_sum, upstream, scale, g) -> None:
    _run("tribe_nce_grad", _ptr(logits), n, logits.stride(0), shift, _ptr(row_sum), _ptr(col_sum), _ptr(upstream), scale, _ptr(g),
                                     g.stride(0), _stream())


# ---------------------------------------------------------------------------------------------- alternative losses
LOSS_SMOOTH_L1, LOSS_HUBER, LOSS_L1 = 1, 2, 3


def point_loss_fwd_bwd(pred, target, kind: int, param: float, want_grad=True, grad_scale=1.0):
    """SmoothL1 / Huber / L1 with reduction="mean": loss[1] fp32 (+ gradient wrt pred in the same pass)."""
    _need(pred, torch.float32, "loss pred"), _need(target, torch.float32, "loss target")
    loss = torch.empty(1, device=pred.device, dtype=torch.float32)
    grad = torch.empty_like(pred) if want_grad else None
    partial = torch.empty(1024, device=pred.device, dtype=torch.float64)
    _run("tribe_point_loss_fwd_bwd", _ptr(pred), _ptr(target), _ptr(loss), _ptr(grad), kind, float(param), grad_scale, pred.numel(),
         _ptr(partial), _stream())
    return loss, grad


def pearson_loss_fwd(pred, target, *, layout: str, reduction_mean: bool = True, want_coef: bool = True):
    """PearsonLoss(dim=1) forward on (N, O) ("no") or (B, O, T) ("bdt") tensors -> (loss[1], coef (4, O) or None)."""
    o = pred.shape[1]
    stats = torch.zeros(1, 6, o, device=pred.device, dtype=torch.float64)
    shift = torch.empty(2, o, device=pred.device, dtype=torch.float32)
    pearson_pick_shift(pred, target, shift, layout=layout)
    pearson_stats(pred, target, stats, layout=layout, shift=shift)
    loss = torch.empty(1, device=pred.device, dtype=torch.float32)
    coef = torch.empty(4, o, device=pred.device, dtype=torch.float32) if want_coef else None
    _run("tribe_pearson_loss_finalize", _ptr(stats), o, int(reduction_mean), _ptr(shift), _ptr(coef), _ptr(loss), _stream())
    return loss, coef


def pearson_loss_bwd(pred, target, coef, upstream, *, layout: str, reduction_mean: bool = True):
    _need(pred, torch.float32, "pearson loss pred"), _need(target, torch.float32, "pearson loss target")
    grad = torch.empty_like(pred)
    t_len = 1 if layout == "no" else pred.shape[2]
    _run("tribe_pearson_loss_bwd", _ptr(pred), _ptr(target), _ptr(coef), _ptr(upstream), int(reduction_mean), _ptr(grad), pred.numel(),
         pred.shape[1], t_len, _stream())
    return grad


# ---------------------------------------------------------------------------------------------- around the path (§8f)
def gather_windows(src_ptrs, src_dtype, t_total, dst_start, src_start, length, out) -> None:
    """out fp32 (B, rows, T): window b <- rows x [src_start[b], +length[b]) of its timeline array, zero padded."""
    _need(out, torch.float32, "gather_windows out")
    n, rows, t = out.shape
    _run("tribe_gather_windows", _ptr(src_ptrs), _DT[src_dtype], _ptr(t_total), _ptr(dst_start), _ptr(src_start), _ptr(length), _ptr(out),
         n, rows, t, _stream())


def ensemble_weights(r, temperature: float, axis: int = 1):
    """r (M, O) -> softmax(r / temperature) over parcels (axis=1, the reference's behaviour) or members (axis=0)."""
    _need(r, torch.float32, "ensemble r")
    w = torch.empty_like(r)
    _run("tribe_ensemble_weights", _ptr(r), r.shape[0], r.shape[1], float(temperature), int(axis), _ptr(w), _stream())
    return w


def ensemble_average(preds, w=None):
    """preds (M, N, O) fp32 stacked member predictions, w (M, O) or None (plain mean) -> (N, O)."""
    _need(preds, torch.float32, "ensemble preds")
    m, n, o = preds.shape
    out = torch.empty(n, o, device=preds.device, dtype=torch.float32)
    _run("tribe_ensemble_average", _ptr(preds), _ptr(w), m, n, o, _ptr(out), _stream())
    return out


def mean_lastdim(x):
    _need(x, torch.float32, "mean_lastdim x")
    y = torch.empty(x.shape[:-1], device=x.device, dtype=torch.float32)
    _run("tribe_mean_lastdim", _ptr(x), _ptr(y), y.numel(), x.shape[-1], _stream())
    return y


def retrieval_ranks(x, y, want_scores=False):
    _need(x, torch.float32, "retrieval x"), _need(y, torch.float32, "retrieval y")
    n, c = x.shape
    if y.shape != x.shape:
        raise TribeError("retrieval_ranks: x and y must have the same (n, c) shape")
    ranks = torch.empty(n, device=x.device, dtype=torch.float32)
    scores = torch.empty(n, n, device=x.device, dtype=torch.float32) if want_scores else None
    _run("tribe_retrieval_ranks", _ptr(x), _ptr(y), n, c, _ptr(ranks), _ptr(scores), _stream())
    return ranks, scores


def swa_update(avg, params, n_averaged: int) -> None:
    _need(avg, torch.float32, "swa avg"), _need(params, torch.float32, "swa params")
    _run("tribe_swa_update", _ptr(avg), _ptr(params), avg.numel(), int(n_averaged), _stream())


def gemm_set_sm_limit(n_sms: int) -> None:
    """Persistent GEMM grids use at most ``n_sms`` SMs from the next launch on (0 = all SMs)."""
    _run("tribe_gemm_set_sm_limit", int(n_sms))


def transpose_last2(x):
    """(B, R, C) fp32 -> contiguous (B, C, R), bit-exact."""
    _need(x, torch.float32, "transpose x")
    b, r, c = x.shape
    y = torch.empty(b, c, r, device=x.device, dtype=torch.float32)
    if x.numel():
        _run("tribe_transpose_last2", _ptr(x), _ptr(y), b, r, c, _stream())
    return y


FUSED_ATTN = bool(int(__import__("os").environ.get("TRIBE_FUSED_ATTN", "1")))


def attn_fusable(T: int, dh: int) -> bool:
    """Whole score rows fit one CTA's TMEM accumulator (<= 320 keys) and head dims tile by 64."""
    return FUSED_ATTN and T <= 320 and dh % 64 == 0


def attn_scores(a, a_off: int, b, b_off: int, n_batch: int, T: int, heads: int, dh: int, scale: float, out, *, p_in=None) -> None:
    """mode 0 (p_in None): out = softmax(scale * a b^T) per (batch, head); mode 1: out = P o (a b^T - rowsum(a b^T o P)) * scale.
    a, b: bf16 (n_batch * T, ld) with head h at columns [off + h*dh, ...); out / p_in: bf16 (n_batch * heads, T, Tp)."""
    _need(a, torch.bfloat16, "attn a"), _need(b, torch.bfloat16, "attn b"), _need(out, torch.bfloat16, "attn out")
    _run("tribe_attn_scores", _ptr(a), a.shape[-1], a_off, _ptr(b), b.shape[-1], b_off, n_batch, T, heads, dh, float(scale),
         0 if p_in is None else 1, _ptr(p_in), _ptr(out), out.shape[-1], _stream())


# opt-in: correct and one launch instead of two, but measured slower than the two-launch path at TRIBE's shape
FUSED_ATTN_FWD = bool(int(__import__("os").environ.get("TRIBE_FUSED_ATTN_FWD", "0")))


def attn_fwd_fusable(T: int, dh: int) -> bool:
    """The flash-style forward (scores + softmax + P.V in one launch) needs two score parts (160 < T <= 320) and a head dim
    that splits into one or two MMA-N halves of 64..256."""
    return FUSED_ATTN and FUSED_ATTN_FWD and attn_fwd_supported(T, dh)


def attn_fwd_supported(T: int, dh: int) -> bool:
    half = dh // 2 if dh > 256 else dh
    return 160 < T <= 320 and dh % 64 == 0 and half % 64 == 0 and half <= 256 and dh <= 448


def attn_fwd(qkv, q_off: int, k_off: int, v_off: int, n_batch: int, T: int, heads: int, dh: int, scale: float, out, out_off: int = 0, *,
             p_out=None) -> None:
    """out[(b, t), out_off + h*dh + :] = softmax(scale * q k^T) v per (batch, head); q / k / v are column blocks of the packed
    bf16 (n_batch * T, ld) buffer ``qkv``; ``p_out`` (optional bf16 (n_batch*heads, T, Tp)) receives P for the backward."""
    _need(qkv, torch.bfloat16, "attn qkv"), _need(out, torch.bfloat16, "attn out")
    if p_out is not None:
        _need(p_out, torch.bfloat16, "attn P")
    Tp = p_out.shape[-1] if p_out is not None else (T + 7) // 8 * 8
    _run("tribe_attn_fwd", _ptr(qkv), qkv.shape[-1], q_off, _ptr(qkv), qkv.shape[-1], k_off, _ptr(qkv), qkv.shape[-1], v_off, n_batch, T, heads, dh,
         float(scale), _ptr(p_out), Tp, _ptr(out), out.shape[-1], out_off, _stream())


def zero_(t) -> None:
    """Stream-ordered zero fill of a contiguous CUDA tensor (cudaMemsetAsync through the C ABI; capturable)."""
    if not t.is_contiguous():
        raise TribeError("zero_: tensor must be contiguous")
    if t.numel():
        _run("tribe_memset_zero", _ptr(t), t.numel() * t.element_size(), _stream())


def copy_(dst, src) -> None:
    """Stream-ordered copy-engine copy between contiguous CUDA tensors of equal byte size (cudaMemcpyAsync through the ABI)."""
    if not (dst.is_cuda and src.is_cuda and dst.is_contiguous() and src.is_contiguous()) or dst.numel() * dst.element_size() != src.numel() * src.element_size():
        raise TribeError("copy_: contiguous CUDA tensors of equal byte size required")
    if dst.numel():
        _run("tribe_memcpy_async", _ptr(dst), _ptr(src), dst.numel() * dst.element_size(), _stream())


def copy_to_pinned(dst_host, src) -> None:
    """Stream-ordered device -> PINNED host copy through the C ABI (cudaMemcpyAsync, capturable).  Deliberately not
    ``dst.copy_(src, non_blocking=True)``: torch's caching host allocator then tracks the pinned block's stream uses with
    events, and an event recorded while the stream was being captured makes a later, unrelated host allocation fail once
    with "invalid argument" (seen as a flaky ``.item()`` after graph captures)."""
    if dst_host.is_cuda or not dst_host.is_pinned() or not src.is_cuda or not (dst_host.is_contiguous() and src.is_contiguous()):
        raise TribeError("copy_to_pinned: contiguous pinned host destination and CUDA source required")
    if dst_host.numel() * dst_host.element_size() != src.numel() * src.element_size():
        raise TribeError("copy_to_pinned: sizes differ")
    if src.numel():
        _run("tribe_memcpy_async", ctypes.c_void_p(dst_host.data_ptr()), _ptr(src), src.numel() * src.element_size(), _stream())


def scale_dev(src, scalar):
    """src * scalar with ``scalar`` a one-element CUDA tensor (no host read)."""
    _need(src, torch.float32, "scale src")
    s = scalar.detach().to(src.device, torch.float32).reshape(1).contiguous()
    out = torch.empty_like(src)
    _run("tribe_scale_dev", _ptr(src), _ptr(s), _ptr(out), src.numel(), _stream())
    return out
