"""CPU model of the half2 exponential of the streaming kernel's steady state (healnet_b200/csrc/xattn_small.cu:
ex2_pair_h2_lean), instruction by instruction in numpy fp16 / integer arithmetic, checked EXHAUSTIVELY over every fp16
input pattern: the bit tricks it rests on are easy to get subtly wrong, and on the device they are only exercised
through the parity tests.

Device sequence per packed pair (x0, x1):  F2FP (fp32 -> fp16x2, rn) ; VIMNMX.U16x2 min with 0xCC00 (clamps negative
values at -16: negative fp16 patterns order by magnitude) ; t = x + 1552 (HADD2: the integer rint(x) + 16 lands in the
low mantissa bits, t's pattern is 0x6600 + n') ; f = x - (t - 1552) ; degree-3 Horner in HFMA2 ; r = t_bits * 1024 +
p_bits as ONE 32-bit multiply-add ; lane-wise add of 0xBE68C000 fused with a signed 16-bit max against 0.
The caller ORs the t words of all polynomial lanes: a bit outside 0x661F means "some x rounds to 16 or more" (P above
2^15, the raise threshold) and sends the tile to the exact path.

What is checked: (1) no carry crosses the lanes and the constant the low lane pushes into the high lane is 0x198 for
every admissible input; (2) the result is 2^x to ~2e-3 (fp16 arithmetic) for -14 <= x < 15.5, an under-estimate (never more than the true
value) in the subnormal band -15.5 < x < -14 — weights below 2^-24 of the reference — and an exact 0 for x <= -15.5
(exponent through zero -> negative lane -> flushed): never a spurious positive weight; (3) the overflow
test fires exactly for the inputs that round to >= 16 (incl. +inf and positive NaN) and never otherwise."""
import numpy as np

POLY_T_OK = 0x661F
C3, C2, C1, C0 = 0.05517167, 0.24261112, 0.69326099, 0.99992807


def _h(a):
    return np.asarray(a, dtype=np.float64).astype(np.float16)


def lanes(xh):
    """xh: fp16 array (what F2FP delivered) -> (t bits, p bits) per lane, uint16."""
    bits = np.minimum(xh.view(np.uint16), np.uint16(0xCC00))          # min.u16x2 with bits(-16)
    x = bits.view(np.float16).astype(np.float64)
    with np.errstate(over="ignore", invalid="ignore"):
        t = _h(x + 1552.0)                                             # HADD2 (one rounding)
        n = _h(t.astype(np.float64) - 1552.0)
        f = _h(x - n.astype(np.float64)).astype(np.float64)
        p = _h(np.full_like(f, np.float64(np.float16(C3))))
        for c in (C2, C1, C0):                                         # HFMA2: fused, one rounding
            p = _h(p.astype(np.float64) * f + np.float64(np.float16(c)))
    return t.view(np.uint16), p.view(np.uint16)


def combine(t_lo, p_lo, t_hi, p_hi):
    """The integer tail on one 32-bit register holding two lanes -> (result lo, result hi) as uint16."""
    T = t_lo.astype(np.uint64) | (t_hi.astype(np.uint64) << np.uint64(16))
    P = p_lo.astype(np.uint64) | (p_hi.astype(np.uint64) << np.uint64(16))
    r = (T * np.uint64(1024) + P) & np.uint64(0xFFFFFFFF)              # IMAD, 32-bit wrap
    lo = ((r & np.uint64(0xFFFF)) + np.uint64(0xC000)) & np.uint64(0xFFFF)   # add.u16x2 with 0xBE68C000
    hi = ((r >> np.uint64(16)) + np.uint64(0xBE68)) & np.uint64(0xFFFF)
    flush = lambda v: np.where(v >= 0x8000, 0, v).astype(np.uint16)   # max.s16x2 against 0
    return flush(lo), flush(hi)


def _all_fp16():
    return np.arange(65536, dtype=np.uint32).astype(np.uint16).view(np.float16)


def test_lean_half2_exponential_is_exact_enough_and_never_spurious():
    xh = _all_fp16()
    t, p = lanes(xh)
    x = xh.astype(np.float64)
    ok_in = np.isfinite(x) & (x < 15.5)          # inputs the steady state may keep (the others must raise, see below)
    for partner in (np.float16(-16.0), np.float16(0.0), np.float16(15.0)):   # the OTHER lane of the pair, in range
        tp, pp = lanes(np.full_like(xh, partner))
        lo, _ = combine(t, p, tp, pp)             # value under test in the low lane
        _, hi = combine(tp, pp, t, p)             # ... and in the high lane
        want_partner = 2.0 ** float(partner) if partner > -15.5 else 0.0
        for got in (lo, hi):
            g = got.view(np.float16).astype(np.float64)
            flushed = ok_in & (x <= -15.5)
            assert np.all(g[flushed] == 0.0), "a weight below 2^-15.5 of the reference must flush to an exact zero"
            live = ok_in & (x >= -14.0)           # normal fp16 results (exponent field >= 1)
            rel = np.abs(g[live] / 2.0 ** x[live] - 1.0)
            assert rel.max() < 2.5e-3, rel.max()  # fp16 Horner + fp16 argument (the kernel's P is fp16 anyway)
            # -15.5 < x < -14: the exponent field is 0, the lane reads as a subnormal: a weight below 2^-24 of the
            # reference that is under-estimated (by up to 40 %), never over-estimated and never negative
            band = ok_in & (x > -15.5) & (x < -14.0)
            assert np.all(g[band] >= 0.0) and np.all(g[band] <= 2.0 ** x[band] * (1 + 2.5e-3))
        # the in-range partner lane is untouched by whatever the lane under test holds (no carry, constant spill-over)
        lo_p, hi_p = combine(tp, pp, t, p)[0], combine(t, p, tp, pp)[1]
        for got in (lo_p[ok_in], hi_p[ok_in]):
            g = got.view(np.float16).astype(np.float64)
            assert np.allclose(g, want_partner, rtol=2.5e-3, atol=0.0)


def test_overflow_shows_in_the_or_of_the_t_words_and_only_then():
    xh = _all_fp16()
    t, _ = lanes(xh)
    x = xh.astype(np.float64)
    fires = (t & np.uint16(~POLY_T_OK & 0xFFFF)) != 0
    # after the unsigned clamp every NEGATIVE pattern (incl. -inf, negative NaNs) is -16 or above: never fires
    neg = xh.view(np.uint16) >= 0x8000
    assert not fires[neg].any()
    pos = ~neg
    must = pos & (np.isnan(x) | (x >= 15.5))      # rounds to 16 or more (ties to even: 15.5 -> 16), +inf, +NaN
    assert fires[must].all()
    assert not fires[pos & ~must].any()
    # OR-accumulation loses nothing: a word in range only has bits inside 0x661F
    in_range = pos & ~must
    assert np.all((t[in_range] & np.uint16(~POLY_T_OK & 0xFFFF)) == 0)
    assert np.all((t[in_range] >= 0x6600) & (t[in_range] <= 0x661F))
    assert np.all(t[neg] >= 0x6600) and np.all(t[neg] <= 0x6610)
