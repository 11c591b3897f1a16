"""TEST INFRASTRUCTURE — CPU restatement of the third-party ``x_transformers.Encoder``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this file.  The product path never does.

Why a restatement: the reference builds its fusion encoder with
``x_transformers.Encoder(dim=3072, heads=8, depth=8, attn_dim_head=384, use_scalenorm=True,
rotary_pos_emb=True, scale_residual=True, attn_flash=False, ff_mult=4, ...)``
(reference ``modeling_utils/modeling_utils/models/transformer.py:43-61``, called from
``algonauts2025/model.py:109-111`` and run at ``model.py:173``).  ``x_transformers`` is declared as
``x_transformers>=1.27.20`` (``modeling_utils/pyproject.toml:12``), is not vendored under /root/reference, has no
lock file and is not installed in this image.  **PARITY UNPINNED** for this block: the reference ships no test or golden
vector for it, so the semantics below are *declared* (x_transformers >= 1.30 / 2.x behaviour):

* ``ScaleNorm``: ``y = F.normalize(x, dim=-1, eps=1e-12) * sqrt(dim) * g`` with ``g = ones(1)``.
* layer pattern ``('a', 'f') * depth``, pre-norm, one final ScaleNorm, no masks, no dropout.
* ``Residual(scale_residual=True)``: ``out = branch(x_normed) + residual * residual_scale`` (``ones(dim)``).
* ``Attention``: bias-free ``to_q/to_k/to_v/to_out``; heads split ``b n (h d) -> b h n d``; rotary on the first
  ``rot_dim = max(dim_head // 2, 32)`` dims of every head of q and k (v untouched), *interleaved* pairs
  ``(0,1),(2,3),...``, ``inv_freq = 10000 ** (-arange(0, rot_dim, 2) / rot_dim)``, angles computed in fp32;
  ``sim = q k^T * dim_head ** -0.5``; ``softmax(dtype=float32)``.
* ``FeedForward``: ``Linear(dim, 4 dim, bias) -> GELU(erf) -> Linear(4 dim, dim, bias)``.
* ``semantics="v1.27"`` (the oldest release the pin admits, x_transformers 1.27.x) differs in exactly two places:
  ``ScaleNorm`` is ``x / norm(x).clamp(min=1e-5) * g`` with ``g = ones(1) * dim ** -0.5`` (no ``sqrt(dim)`` factor), and the
  rotary embedding is *half-split*: ``freqs = cat((f, f))``, ``rotate_half`` pairs dim ``i`` with ``i + rot_dim/2``.
  A checkpoint trained under one release must be evaluated with that release's arithmetic (the state-dict keys are
  identical), hence the switch.
* parameter names mirror the library's module nesting so that a reference checkpoint's ``state_dict`` keys line up:
  ``layers.<i>.0.0.g``, ``layers.<i>.1.{to_q,to_k,to_v,to_out}.weight``, ``layers.<i>.1.ff.0.0.{weight,bias}``,
  ``layers.<i>.1.ff.2.{weight,bias}``, ``layers.<i>.2.residual_scale``, ``final_norm.g``,
  buffer ``rotary_pos_emb.inv_freq``.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn


SEMANTICS = ("v2", "v1.27")


class ScaleNorm(nn.Module):
    def __init__(self, dim: int, semantics: str = "v2"):
        super().__init__()
        self.semantics = semantics
        if semantics == "v1.27":
            self.eps = 1e-5
            self.g = nn.Parameter(torch.ones(1) * (dim**-0.5))
        else:
            self.scale = dim**0.5
            self.g = nn.Parameter(torch.ones(1))

    def forward(self, x):
        if self.semantics == "v1.27":
            norm = torch.norm(x, dim=-1, keepdim=True)
            return x / norm.clamp(min=self.eps) * self.g
        return F.normalize(x, dim=-1) * self.scale * self.g


class RotaryEmbedding(nn.Module):
    def __init__(self, dim: int, base: float = 10000.0, semantics: str = "v2"):
        super().__init__()
        self.semantics = semantics
        inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2).float() / dim))
        self.register_buffer("inv_freq", inv_freq)

    def forward(self, seq_len: int, device=None):
        t = torch.arange(seq_len, device=device).type_as(self.inv_freq)
        freqs = torch.einsum("i,j->ij", t, self.inv_freq)
        if self.semantics == "v1.27":
            return torch.cat((freqs, freqs), dim=-1)  # half-split: (f0, f1, ..., f0, f1, ...)
        # interleave: (f0, f0, f1, f1, ...)
        return torch.stack((freqs, freqs), dim=-1).reshape(seq_len, -1)


def _rotate_pairs(x):
    x = x.reshape(*x.shape[:-1], -1, 2)
    x1, x2 = x.unbind(dim=-1)
    return torch.stack((-x2, x1), dim=-1).reshape(*x.shape[:-2], -1)


def _rotate_half(x):
    x1, x2 = x.reshape(*x.shape[:-1], 2, -1).unbind(dim=-2)
    return torch.cat((-x2, x1), dim=-1)


def apply_rotary(t, freqs, semantics: str = "v2"):
    rot_dim = freqs.shape[-1]
    dtype = t.dtype
    t_rot, t_pass = t[..., :rot_dim], t[..., rot_dim:]
    rot = _rotate_half if semantics == "v1.27" else _rotate_pairs
    t_rot = t_rot.float() * freqs.cos() + rot(t_rot.float()) * freqs.sin()
    return torch.cat((t_rot.to(dtype), t_pass), dim=-1)


class Attention(nn.Module):
    def __init__(self, dim: int, dim_head: int, heads: int, semantics: str = "v2"):
        super().__init__()
        self.heads, self.dim_head, self.semantics = heads, dim_head, semantics
        self.scale = dim_head**-0.5
        inner = dim_head * heads
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_k = nn.Linear(dim, inner, bias=False)
        self.to_v = nn.Linear(dim, inner, bias=False)
        self.to_out = nn.Linear(inner, dim, bias=False)

    def forward(self, x, freqs=None):
        b, n, _ = x.shape
        h, d = self.heads, self.dim_head
        q, k, v = (f(x).view(b, n, h, d).transpose(1, 2) for f in (self.to_q, self.to_k, self.to_v))
        if freqs is not None:
            q, k = apply_rotary(q, freqs, self.semantics), apply_rotary(k, freqs, self.semantics)
        sim = torch.einsum("bhid,bhjd->bhij", q, k) * self.scale
        attn = sim.softmax(dim=-1, dtype=torch.float32).to(sim.dtype)
        out = torch.einsum("bhij,bhjd->bhid", attn, v)
        return self.to_out(out.transpose(1, 2).reshape(b, n, h * d))


class FeedForward(nn.Module):
    def __init__(self, dim: int, mult: int = 4):
        super().__init__()
        inner = int(dim * mult)
        self.ff = nn.Sequential(nn.Sequential(nn.Linear(dim, inner), nn.GELU()), nn.Dropout(0.0), nn.Linear(inner, dim))

    def forward(self, x):
        return self.ff(x)


class Residual(nn.Module):
    def __init__(self, dim: int, scale_residual: bool):
        super().__init__()
        self.residual_scale = nn.Parameter(torch.ones(dim)) if scale_residual else None

    def forward(self, x, residual):
        if self.residual_scale is not None:
            residual = residual * self.residual_scale
        return x + residual


class Encoder(nn.Module):
    """``x_transformers.Encoder`` restatement; accepts (and checks) exactly the kwargs the reference passes."""

    def __init__(self, dim: int, depth: int, heads: int = 8, attn_dim_head: int = 64, ff_mult: int = 4,
                 use_scalenorm: bool = False, rotary_pos_emb: bool = False, scale_residual: bool = False, semantics: str = "v2", **kw):
        super().__init__()
        if semantics not in SEMANTICS:
            raise ValueError(f"semantics must be one of {SEMANTICS}")
        self.semantics = semantics
        unsupported = {"cross_attend": False, "attn_flash": False, "attn_dropout": 0.0, "ff_dropout": 0.0,
                       "use_rmsnorm": False, "rel_pos_bias": False, "alibi_pos_bias": False, "rotary_xpos": False,
                       "residual_attn": False, "layer_dropout": 0.0}
        for key, val in kw.items():
            if key not in unsupported or unsupported[key] != val:
                raise NotImplementedError(f"oracle Encoder restates only the TRIBE configuration; got {key}={val!r}")
        if not use_scalenorm:
            raise NotImplementedError("oracle Encoder restates only use_scalenorm=True")
        self.dim, self.depth, self.heads, self.dim_head = dim, depth, heads, attn_dim_head
        self.rotary_pos_emb = RotaryEmbedding(max(attn_dim_head // 2, 32), semantics=semantics) if rotary_pos_emb else None
        layers = []
        for _ in range(depth):
            for kind in ("a", "f"):
                block = Attention(dim, attn_dim_head, heads, semantics) if kind == "a" else FeedForward(dim, ff_mult)
                norms = nn.ModuleList([ScaleNorm(dim, semantics), None, None])
                layers.append(nn.ModuleList([norms, block, Residual(dim, scale_residual)]))
        self.layers = nn.ModuleList(layers)
        self.final_norm = ScaleNorm(dim, semantics)

    def forward(self, x):
        freqs = self.rotary_pos_emb(x.shape[1], x.device) if self.rotary_pos_emb is not None else None
        for norms, block, residual_fn in self.layers:
            residual = x
            x = norms[0](x)
            out = block(x, freqs) if isinstance(block, Attention) else block(x)
            x = residual_fn(out, residual)
        return self.final_norm(x)


class Decoder(Encoder):  # imported (never built) by the reference: transformer.py:44
    def __init__(self, *a, **k):
        raise NotImplementedError("causal Decoder is not on the TRIBE path")


def param_count(dim=3072, depth=8, ff_mult=4):
    per = 4 * dim * dim + 2 * dim * dim * ff_mult + dim * ff_mult + dim + 2 + 2 * dim
    return per * depth + 1


assert param_count() == 906_141_713
assert math.isclose(384**-0.5, 0.05103103630798288)
