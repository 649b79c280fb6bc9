"""CPU oracle for the HEALNet fusion forward — TEST INFRASTRUCTURE ONLY.

A functional restatement (torch CPU tensors, fp32 or fp64, no nn.Module) of the reference's
``healnet/models/healnet.py`` forward pass, driven purely by a reference-format ``state_dict``.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this module; the product (``healnet_b200``) never does and has no CPU fallback.

Parity pin: the reference has no golden vectors of its own (its tests assert shapes only,
healnet/tests/test_healnet.py:26-67), so this restatement is pinned against *outputs of the
unmodified reference executed in the build container* — ``tests/golden/make_golden.py`` imports
/root/reference/healnet/models/healnet.py by path and stores inputs, state_dicts and outputs under
``tests/golden/*.npz``; ``tests/test_oracle.py`` checks this file against them (max abs ~1e-6).

Each function cites the reference lines it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F


@dataclass
class OracleConfig:
    """Constructor hyper-parameters of the reference (healnet/models/healnet.py:15-38)."""

    n_modalities: int
    channel_dims: Sequence[int]
    num_spatial_axes: Sequence[int]
    out_dims: int
    depth: int = 3
    num_freq_bands: int = 2
    max_freq: float = 10.0
    l_c: int = 128
    l_d: int = 128
    x_heads: int = 8
    l_heads: int = 8
    cross_dim_head: int = 64
    latent_dim_head: int = 64
    fourier_encode_data: bool = True
    self_per_cross_attn: int = 1
    final_classifier_head: bool = True
    snn: bool = True
    extra: dict = field(default_factory=dict)


def fourier_table(size: int, max_freq: float, num_bands: int, dtype=torch.float32, device=None) -> torch.Tensor:
    """Per-axis feature table, rows = positions, cols = [sin(pi p f_k)]_k ++ [cos(pi p f_k)]_k ++ [p].

    healnet.py:212 (linspace(-1,1,size)), :292-302 (fourier_encode: scales = linspace(1, max_freq/2, B),
    x*scales*pi, cat(sin, cos), cat(.., orig_x)).
    """
    p = torch.linspace(-1.0, 1.0, steps=size, dtype=dtype, device=device)  # (the reference builds it on data.device, :212)
    scales = torch.linspace(1.0, max_freq / 2, num_bands, dtype=dtype, device=device)
    x = p[:, None] * scales[None, :] * math.pi
    return torch.cat([x.sin(), x.cos(), p[:, None]], dim=-1)


def encode_modality(data: torch.Tensor, n_axes: int, max_freq: float, num_bands: int, fourier: bool) -> torch.Tensor:
    """(b, *axes, c) -> (b, N, c + A(2B+1)) with N = prod(axes) (row-major). healnet.py:205-222."""
    b, *axes, c = data.shape
    assert len(axes) == n_axes, "input data must have the declared number of axes"  # healnet.py:207-208
    if fourier:
        feats = []
        for a, size in enumerate(axes):  # meshgrid(indexing='ij') + '... n d -> ... (n d)'  healnet.py:213-215
            t = fourier_table(size, max_freq, num_bands, data.dtype, data.device)  # (size, 2B+1)
            shape = [1] * len(axes) + [t.shape[1]]
            shape[a] = size
            feats.append(t.reshape(shape).expand(*axes, t.shape[1]))
        enc = torch.cat(feats, dim=-1)  # (*axes, A(2B+1)): axis-major feature order
        data = torch.cat((data, enc.unsqueeze(0).expand(b, *enc.shape)), dim=-1)  # healnet.py:216-217
    return data.reshape(b, -1, data.shape[-1])  # healnet.py:221


def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """nn.LayerNorm over the last dim, eps 1e-5, biased variance (healnet.py:310-311)."""
    return F.layer_norm(x, (x.shape[-1],), w.to(x.dtype), b.to(x.dtype), 1e-5)


def attention(x: torch.Tensor, ctx: torch.Tensor, wq, wkv, wo, bo, heads: int,
              mask: Optional[torch.Tensor] = None, head_chunk: int = 0,
              want_weights: bool = False):
    """Attention.forward, healnet.py:400-426. x (b,L,D) normed latents, ctx (b,N,C) normed context.

    softmax((q k^T * dh^-0.5) / 0.5) (scale :375,:409; temperature 0.5 :419, :354-365); masked keys get
    -finfo.max before the temperature division (:411-415); to_out = Linear + LeakyReLU(0.01) (:383-386).
    ``head_chunk`` > 0 evaluates that many heads at a time to bound memory; results are identical up to
    fp reassociation inside torch's matmul.
    """
    b, L, _ = x.shape
    N = ctx.shape[1]
    inner = wq.shape[0]
    dh = inner // heads
    q = x @ wq.to(x.dtype).t()
    kv = ctx @ wkv.to(x.dtype).t()
    k, v = kv[..., :inner], kv[..., inner:]  # chunk(2) healnet.py:405
    q = q.reshape(b, L, heads, dh).permute(0, 2, 1, 3)  # '(b h) n d' healnet.py:407
    k = k.reshape(b, N, heads, dh).permute(0, 2, 1, 3)
    v = v.reshape(b, N, heads, dh).permute(0, 2, 1, 3)
    out = torch.empty(b, heads, L, dh, dtype=x.dtype, device=x.device)
    weights = [] if want_weights else None
    step = head_chunk if head_chunk > 0 else heads
    for h0 in range(0, heads, step):
        sl = slice(h0, min(heads, h0 + step))
        sim = torch.matmul(q[:, sl], k[:, sl].transpose(-1, -2)) * (dh ** -0.5)
        if mask is not None:
            m = mask.reshape(b, -1)[:, None, None, :]
            sim = sim.masked_fill(~m, -torch.finfo(sim.dtype).max)
        attn = torch.softmax(sim / 0.5, dim=-1)
        if want_weights:
            weights.append(attn)
        out[:, sl] = torch.matmul(attn, v[:, sl])
    out = out.permute(0, 2, 1, 3).reshape(b, L, inner)  # '(b h) n d -> b n (h d)' healnet.py:425
    y = F.leaky_relu(out @ wo.to(x.dtype).t() + bo.to(x.dtype), 0.01)
    if want_weights:
        return y, torch.cat(weights, dim=1).reshape(b * heads, L, N)
    return y


def feed_forward(x: torch.Tensor, w1, b1, w2, b2, snn: bool) -> torch.Tensor:
    """FeedForward, healnet.py:339-351; gate = a * selu(g) (:328-331) or a * gelu(g) (:323-326)."""
    h = x @ w1.to(x.dtype).t() + b1.to(x.dtype)
    a, g = h.chunk(2, dim=-1)
    h = a * (F.selu(g) if snn else F.gelu(g))
    return h @ w2.to(x.dtype).t() + b2.to(x.dtype)


def forward(sd: Dict[str, torch.Tensor], cfg: OracleConfig, tensors: List[Optional[torch.Tensor]],
            mask: Optional[torch.Tensor] = None, return_embeddings: bool = False, verbose: bool = False,
            dtype=torch.float32, head_chunk: int = 0, collect_weights: Optional[list] = None) -> torch.Tensor:
    """HealNet.forward, healnet.py:190-250, default ``verbose=False`` semantics.

    A modality that is ``None`` (or absent because the list is short) has its cross-attention + cross-FF
    skipped (the reference raises inside ``try`` and swallows, :235-239) while the latent self-attention
    block still runs (:241-245). With ``verbose=True`` a modality passed as ``None`` skips the latent block as
    well (the ``continue`` at :229-232 is nested under ``if verbose``). The caller's list is not mutated (the
    reference does, :222).
    """
    M = cfg.n_modalities
    ctxs: List[Optional[torch.Tensor]] = []
    b = None
    for i in range(M):
        t = tensors[i] if i < len(tensors) else None
        if t is None:
            ctxs.append(None)
            continue
        t = t.to(dtype)
        b = t.shape[0]
        ctxs.append(encode_modality(t, cfg.num_spatial_axes[i], cfg.max_freq, cfg.num_freq_bands,
                                    cfg.fourier_encode_data))
    assert b is not None, "at least one modality is required"
    g = lambda k: sd[k].to(dtype)
    x = g("latents").unsqueeze(0).expand(b, -1, -1).clone()  # healnet.py:225
    for l in range(cfg.depth):  # healnet.py:227
        for i in range(M):  # healnet.py:228
            if verbose and i < len(tensors) and tensors[i] is None:  # :229-232
                continue
            if ctxs[i] is not None:
                p = f"layers.{l}.{2 * i}"
                xn = layer_norm(x, g(p + ".norm.weight"), g(p + ".norm.bias"))  # PreNorm :313-321
                cn = layer_norm(ctxs[i], g(p + ".norm_context.weight"), g(p + ".norm_context.bias"))
                r = attention(xn, cn, g(p + ".fn.to_q.weight"), g(p + ".fn.to_kv.weight"),
                              g(p + ".fn.to_out.0.weight"), g(p + ".fn.to_out.0.bias"), cfg.x_heads, mask,
                              head_chunk, want_weights=collect_weights is not None)
                if collect_weights is not None:
                    r, w = r
                    collect_weights.append(w)
                x = r + x  # :236
                p = f"layers.{l}.{2 * i + 1}"
                xn = layer_norm(x, g(p + ".norm.weight"), g(p + ".norm.bias"))
                x = feed_forward(xn, g(p + ".fn.net.0.weight"), g(p + ".fn.net.0.bias"),
                                 g(p + ".fn.net.2.weight"), g(p + ".fn.net.2.bias"), cfg.snn) + x  # :237
            if cfg.self_per_cross_attn > 0:  # :241-245 (inside the modality loop)
                p = f"layers.{l}.{2 * M}.0"
                xn = layer_norm(x, g(p + ".norm.weight"), g(p + ".norm.bias"))
                r = attention(xn, xn, g(p + ".fn.to_q.weight"), g(p + ".fn.to_kv.weight"),
                              g(p + ".fn.to_out.0.weight"), g(p + ".fn.to_out.0.bias"), cfg.l_heads, None,
                              head_chunk, want_weights=collect_weights is not None)
                if collect_weights is not None:
                    r, w = r
                    collect_weights.append(w)
                x = r + x
                p = f"layers.{l}.{2 * M}.1"
                xn = layer_norm(x, g(p + ".norm.weight"), g(p + ".norm.bias"))
                x = feed_forward(xn, g(p + ".fn.net.0.weight"), g(p + ".fn.net.0.bias"),
                                 g(p + ".fn.net.2.weight"), g(p + ".fn.net.2.bias"), cfg.snn) + x
    if return_embeddings:  # :247-248
        return x
    if not cfg.final_classifier_head:  # nn.Identity :185
        return x
    pooled = x.mean(dim=1)  # Reduce('b n d -> b d', 'mean') :182
    pooled = layer_norm(pooled, g("to_logits.1.weight"), g("to_logits.1.bias"))
    return pooled @ g("to_logits.2.weight").t() + g("to_logits.2.bias")  # :184


# --------------------------------------------------------------------------- work model (roofline numerators)
def modality_tokens(cfg: OracleConfig, shapes: Sequence[Sequence[int]]):
    """[(N_m, C_m)] for spatial shapes ``shapes[m]`` (without batch and channel dims)."""
    out = []
    for m in range(cfg.n_modalities):
        n = 1
        for s in shapes[m]:
            n *= s
        c = cfg.channel_dims[m] + (cfg.num_spatial_axes[m] * (2 * cfg.num_freq_bands + 1)
                                   if cfg.fourier_encode_data else 0)
        out.append((n, c))
    return out


def flops_per_sample(cfg: OracleConfig, shapes) -> float:
    """As-written algorithmic FLOPs of one forward for one sample."""
    L, D = cfg.l_c, cfg.l_d
    I = cfg.x_heads * cfg.cross_dim_head
    lI = cfg.l_heads * cfg.latent_dim_head
    spc = 1 if cfg.self_per_cross_attn > 0 else 0
    tot = 0.0
    for (n, c) in modality_tokens(cfg, shapes):
        tot += 4 * n * c * I + 2 * L * D * I + 4 * L * n * I + 2 * L * I * D + 24 * L * D * D
        tot += spc * (6 * L * D * lI + 4 * L * L * lI + 2 * L * lI * D + 24 * L * D * D)
    return cfg.depth * tot + 2 * D * cfg.out_dims
