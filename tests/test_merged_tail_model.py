"""CPU model of the streaming kernel's score products (healnet_b200/csrc/xattn_small.cu): the three-term split product
S = Q'h.zh + Q'l.zh + Q'h.zl over 32-column tiles, issued only over the 16-column steps that hold context columns, and
for 17 <= C <= 23 with the MERGED TAIL: columns 16..C-1 of both lo-order products run as ONE 16-column step on
re-arranged tiles —
    z lo tile   [zl_0..15 | zl_16..C-1, 0 (column C), zh_16..C-1, 0..]      (written by the context-row builder)
    Q' lo tile  [Q'l_0..15 | Q'h_16..C-1, 0,           Q'l_16..C-1, 0..]    (re-arranged in shared memory by the row owners)
The model checks, in exact arithmetic on fp16-representable operands, that every mode's UMMA sequence adds up to the
full three-term product."""
import numpy as np
import pytest


def _tiles(C, rng, rows=8, toks=12):
    kd = 32
    q = np.zeros((rows, kd)); z = np.zeros((toks, kd))
    q[:, :C] = rng.standard_normal((rows, C)) * 3
    z[:, :C] = rng.standard_normal((toks, C))
    q[:, C] = rng.standard_normal(rows) * 5          # the folded -m_ref + P_SHIFT
    z[:, C] = 1.0                                    # ones column
    split = lambda a: (a.astype(np.float16).astype(np.float64), (a - a.astype(np.float16).astype(np.float64)).astype(np.float16).astype(np.float64))
    qh, ql = split(q); zh, zl = split(z)
    ql[:, C] = 0.0                                   # the fold is exactly representable in fp16 (the kernel defines it so)
    qh[:, C] = q[:, C].astype(np.float16)
    return qh, ql, zh, zl


def _step(a, b, k):                                  # one 16-column UMMA step: A[:, 16k:16k+16] . B[:, 16k:16k+16]^T
    return a[:, 16 * k:16 * k + 16] @ b[:, 16 * k:16 * k + 16].T


@pytest.mark.parametrize("C", list(range(1, 32)))
def test_issued_steps_add_up_to_the_three_term_product(C):
    rng = np.random.default_rng(C)
    qh, ql, zh, zl = _tiles(C, rng)
    want = qh @ zh.T + ql @ zh.T + qh @ zl.T
    merged = 17 <= C <= 23
    if C <= 15:
        kh, kl = 1, 1
    elif C == 16:
        kh, kl = 2, 1
    elif merged:
        kh, kl = 2, 1
    else:
        kh, kl = 2, 2
    got = sum(_step(qh, zh, k) for k in range(kh))
    got = got + sum(_step(ql, zh, k) for k in range(kl)) + sum(_step(qh, zl, k) for k in range(kl))
    if merged:
        e = C - 16
        zl_t, ql_t = zl.copy(), ql.copy()
        zl_t[:, C + 1:C + 1 + e] = zh[:, 16:C]                       # rowops.cu: hi parts behind the zero column C
        assert np.all(zl_t[:, C] == 0) and np.all(zl_t[:, C + 1 + e:] == 0)
        ql_t[:, 16:32] = 0.0                                          # xattn_small.cu: the row owners' re-arrangement
        ql_t[:, 16:C] = qh[:, 16:C]
        ql_t[:, C + 1:C + 1 + e] = ql[:, 16:C]
        got = got + _step(ql_t, zl_t, 1)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)
