"""CPU: the arithmetic the streaming BACKWARD of the small-context cross-attention will execute (SURVEY.md section 8
row f2; design in DESIGN.md section 9), checked against torch.autograd of the attention exactly as the reference writes
it (healnet/models/healnet.py:403-426 with the context PreNorm :313-321). Nothing of size N is saved by the forward:
the backward needs, per (head, latent row), the merged u = sum_t p_t z_t, the softmax statistics and dO; it streams the
standardised context rows z once more, recomputes p from Q'.z, and accumulates dQ' = sum_t p_t (du.z_t - du.u) z_t.
Test infrastructure only (float64, dense): the product does not import it."""
import math

import torch

torch.manual_seed(0)


def _as_written(xn, ctx, gamma, beta, Wq, Wkv, heads):
    """PreNorm(context) + Attention up to the head merge: returns O (L, heads*dh)."""
    C = ctx.shape[-1]
    cn = torch.nn.functional.layer_norm(ctx, (C,), gamma, beta, 1e-5)
    inner = Wq.shape[0]
    dh = inner // heads
    q = xn @ Wq.t()
    kv = cn @ Wkv.t()
    k, v = kv[:, :inner], kv[:, inner:]
    outs = []
    for h in range(heads):
        sl = slice(h * dh, (h + 1) * dh)
        sim = (q[:, sl] @ k[:, sl].t()) * dh ** -0.5
        p = torch.softmax(sim / 0.5, dim=-1)          # temperature 0.5 (healnet.py:354-365,419)
        outs.append(p @ v[:, sl])
    return torch.cat(outs, dim=-1)


def test_streaming_backward_formulas_match_autograd():
    L, N, D, C, heads, dh = 7, 50, 12, 9, 2, 4
    inner = heads * dh
    dt = torch.float64
    xn = torch.randn(L, D, dtype=dt, requires_grad=True)
    ctx = torch.randn(N, C, dtype=dt) * 2 + 0.5                      # inputs: no gradient flows into the context
    gamma = (1 + 0.3 * torch.randn(C, dtype=dt)).requires_grad_()
    beta = (0.3 * torch.randn(C, dtype=dt)).requires_grad_()
    Wq = (torch.randn(inner, D, dtype=dt) / math.sqrt(D)).requires_grad_()
    Wkv = (torch.randn(2 * inner, C, dtype=dt) / math.sqrt(C)).requires_grad_()
    O = _as_written(xn, ctx, gamma, beta, Wq, Wkv, heads)
    dO = torch.randn_like(O)
    (O * dO).sum().backward()

    # ---- what the kernels do -------------------------------------------------------------------------------
    with torch.no_grad():
        z = (ctx - ctx.mean(-1, keepdim=True)) / torch.sqrt(ctx.var(-1, unbiased=False, keepdim=True) + 1e-5)
        alpha = 2.0 / math.sqrt(dh)
        q = xn @ Wq.t()
        d_xn = torch.zeros_like(xn)
        d_Wq = torch.zeros_like(Wq)
        d_Wkv = torch.zeros_like(Wkv)
        d_gamma = torch.zeros_like(gamma)
        d_beta = torch.zeros_like(beta)
        for h in range(heads):
            sl = slice(h * dh, (h + 1) * dh)
            Wk, Wv = Wkv[sl], Wkv[inner + h * dh: inner + (h + 1) * dh]
            qh, dOh = q[:, sl], dO[:, sl]
            # forward (reassociated): Q' = alpha * gamma * (Wk^T q); p = softmax(Q' z^T); u = p z
            Qp = alpha * gamma * (qh @ Wk)                                # (L, C)
            p = torch.softmax(Qp @ z.t(), dim=-1)                          # the beta.k term is constant per row
            u = p @ z                                                      # (L, C) — all the forward keeps
            # V side: o = Wv (gamma*u + beta)
            g = dOh @ Wv                                                   # (L, C) = Wv^T dO
            d_Wkv[inner + h * dh: inner + (h + 1) * dh] += dOh.t() @ (gamma * u + beta)
            d_gamma += (g * u).sum(0)
            d_beta += g.sum(0)
            du = gamma * g
            # second streaming pass: dS = p * (du.z_t - du.u);  dQ' = dS z
            dS = p * (du @ z.t() - (du * u).sum(-1, keepdim=True))
            dQp = dS @ z                                                   # (L, C)
            # K side: Q' = alpha * gamma * (Wk^T q)
            r = alpha * gamma * dQp                                        # d(Wk^T q)
            d_gamma += (alpha * (qh @ Wk) * dQp).sum(0)
            d_Wkv[sl] += qh.t() @ r                                        # dWk = q r^T
            dq = r @ Wk.t()                                                # (L, dh)
            d_Wq[sl] += dq.t() @ xn
            d_xn += dq @ Wq[sl]
    tol = dict(rtol=1e-9, atol=1e-11)
    torch.testing.assert_close(d_xn, xn.grad, **tol)
    torch.testing.assert_close(d_Wq, Wq.grad, **tol)
    torch.testing.assert_close(d_Wkv, Wkv.grad, **tol)
    torch.testing.assert_close(d_gamma, gamma.grad, **tol)
    torch.testing.assert_close(d_beta, beta.grad, **tol)


def test_streaming_backward_is_additive_over_token_splits():
    """The second pass splits over the token axis like the forward: with the GLOBAL row statistics (max, denominator,
    u) fixed, dQ' is a plain sum of per-split contributions — no second merge pass is needed."""
    L, N, C = 5, 64, 6
    dt = torch.float64
    Qp = torch.randn(L, C, dtype=dt)
    z = torch.randn(N, C, dtype=dt)
    du = torch.randn(L, C, dtype=dt)
    s = Qp @ z.t()
    m = s.max(-1, keepdim=True).values
    den = torch.exp(s - m).sum(-1, keepdim=True)
    p = torch.exp(s - m) / den
    u = p @ z
    full = (p * (du @ z.t() - (du * u).sum(-1, keepdim=True))) @ z
    parts = torch.zeros_like(full)
    for lo in range(0, N, 16):
        zs = z[lo:lo + 16]
        ps = torch.exp(Qp @ zs.t() - m) / den                              # recomputed from the saved statistics
        parts += (ps * (du @ zs.t() - (du * u).sum(-1, keepdim=True))) @ zs
    torch.testing.assert_close(parts, full, rtol=1e-10, atol=1e-12)
