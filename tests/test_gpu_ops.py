"""GPU (B200): kernel-level parity through the C ABI's hn_op_* entry points against plain torch fp32 / fp64
references of the same op (the oracle covers the composed forward; these isolate each kernel)."""
import ctypes
import math

import pytest
import torch

import healnet_b200
from healnet_b200 import _lib
from oracle import healnet_oracle as O

pytestmark = pytest.mark.gpu


def _lib_stream():
    return healnet_b200.load_library(), torch.cuda.current_stream().cuda_stream


def _split(x, seg):
    """[hi | lo] fp16 rows with 64-aligned segments (the layout the split-precision GEMM consumes)."""
    rows, k = x.shape
    out = torch.zeros(rows, 2 * seg, dtype=torch.float16, device=x.device)
    hi = x.half()
    out[:, :k] = hi
    out[:, seg:seg + k] = (x - hi.float()).half()
    return out


# (the 16384-row cases take the cluster / TMA-multicast path: shared A for the narrow outputs, shared B for N = 4096)
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 72, 200), (2048, 512, 512), (5, 1024, 2005),
                                   (16384, 512, 192), (16384, 4096, 128), (16400, 1536, 64)])
@pytest.mark.parametrize("terms", [1, 2, 3])
def test_gemm_split_precision(M, N, K, terms):
    lib, st = _lib_stream()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
    bias = torch.randn(N, device="cuda", generator=g)
    seg = (K + 63) // 64 * 64
    As, Bs = _split(A, seg), _split(B, seg)
    out = torch.full((M, N), float("nan"), device="cuda")
    rc = lib.hn_op_gemm(As.data_ptr(), Bs.data_ptr(), M, N, K, 2 * seg, 2 * seg, 4, 0, bias.data_ptr(), out.data_ptr(),
                        N, terms, seg, seg, 0, st)
    assert rc == 0, _lib.last_error()
    a_eff = A if terms == 3 else A.half().float()
    b_eff = B if terms >= 2 else B.half().float()
    want = (a_eff.double() @ b_eff.double().t() + bias.double()).float()
    tol = 5e-5 if terms == 3 else 1e-4  # what is left: the dropped lo.lo term and fp32 accumulation order over K
    torch.testing.assert_close(out, want, rtol=tol, atol=tol)
    if terms == 3:  # the point of the split: an fp16-operand product would be ~1e-3 off
        plain = (A.half().float() @ B.half().float().t() + bias)
        assert (plain - want).abs().max() > 5 * (out - want).abs().max()


def test_gemm_epilogues():
    lib, st = _lib_stream()
    M, N, K = 256, 256, 128
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda") / math.sqrt(K)
    bias = torch.randn(N, device="cuda")
    seg = 128
    As, Bs = _split(A, seg), _split(B, seg)
    acc = A @ B.t() + bias
    # residual + LeakyReLU (to_out, healnet.py:383-386,426,236)
    x = torch.randn(M, N, device="cuda")
    x0 = x.clone()
    assert lib.hn_op_gemm(As.data_ptr(), Bs.data_ptr(), M, N, K, 2 * seg, 2 * seg, 3, 0, bias.data_ptr(), x.data_ptr(), N,
                          3, seg, seg, 0, st) == 0
    torch.testing.assert_close(x, x0 + torch.nn.functional.leaky_relu(acc, 0.01), rtol=1e-4, atol=1e-4)
    # gated SELU with interleaved (a, g) rows -> split fp16 output (FeedForward, healnet.py:328-331,350)
    out = torch.zeros(M, 2 * 128, dtype=torch.float16, device="cuda")
    assert lib.hn_op_gemm(As.data_ptr(), Bs.data_ptr(), M, N, K, 2 * seg, 2 * seg, 1, 0, bias.data_ptr(), out.data_ptr(),
                          256, 3, seg, seg, 128, st) == 0
    want = acc[:, 0::2] * torch.nn.functional.selu(acc[:, 1::2])
    got = out[:, :128].float() + out[:, 128:].float()
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)
    # GELU gate
    assert lib.hn_op_gemm(As.data_ptr(), Bs.data_ptr(), M, N, K, 2 * seg, 2 * seg, 1, 1, bias.data_ptr(), out.data_ptr(),
                          256, 3, seg, seg, 128, st) == 0
    want = acc[:, 0::2] * torch.nn.functional.gelu(acc[:, 1::2])
    torch.testing.assert_close(out[:, :128].float() + out[:, 128:].float(), want, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("D", [32, 119, 512, 1024, 1500])
def test_layernorm_split(D):
    lib, st = _lib_stream()
    rows, seg = 77, (D + 63) // 64 * 64
    x = torch.randn(rows, D, device="cuda") * 3 + 1
    g, b = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
    y = torch.full((rows, 2 * seg), float("nan"), dtype=torch.float16, device="cuda")
    assert lib.hn_op_layernorm_f16(x.data_ptr(), D, g.data_ptr(), b.data_ptr(), y.data_ptr(), 2 * seg, seg, seg, rows, D,
                                   st) == 0
    want = torch.nn.functional.layer_norm(x, (D,), g, b, 1e-5)
    torch.testing.assert_close(y[:, :D].float() + y[:, seg:seg + D].float(), want, rtol=2e-5, atol=2e-5)
    assert bool((y[:, D:seg] == 0).all()) and bool((y[:, seg + D:] == 0).all())  # zero padding feeds the GEMM


@pytest.mark.parametrize("axes,c", [((7,), 5), ((12, 10), 3), ((3, 6, 5), 3), ((1, 1), 2)])
def test_build_context_matches_reference_encoding(axes, c):
    """z = standardised [raw | Fourier features] rows; Fourier layout / linspace semantics of healnet.py:211-217,292-302."""
    lib, st = _lib_stream()
    b, bands, max_freq = 2, 2, 10.0
    raw = torch.rand((b,) + axes + (c,), device="cuda")
    ctx = O.encode_modality(raw.cpu(), len(axes), max_freq, bands, True)
    C, N = ctx.shape[-1], ctx.shape[1]
    want = torch.nn.functional.layer_norm(ctx, (C,))
    sizes = (ctypes.c_int * 4)(*axes)
    tab = torch.empty(sum(axes) * (2 * bands + 1), device="cuda")
    zw = 32 if C <= 31 else 64
    z = torch.full((b, N, zw), float("nan"), dtype=torch.float16, device="cuda")
    assert lib.hn_op_build_context(raw.data_ptr(), z.data_ptr(), zw, 1, b, c, len(axes), sizes, bands, max_freq, 1,
                                   tab.data_ptr(), st) == 0
    torch.testing.assert_close(z[..., :C].float().cpu(), want, rtol=2e-3, atol=2e-3)  # fp16 storage
    assert bool((z[..., C] == 1).all()) and bool((z[..., C + 1:] == 0).all())
    # split rows [hi | lo] (what the forward streams): hi + lo reproduces the fp32 standardisation
    zs = torch.full((b, N, 2 * zw), float("nan"), dtype=torch.float16, device="cuda")
    assert lib.hn_op_build_context(raw.data_ptr(), zs.data_ptr(), zw, 2, b, c, len(axes), sizes, bands, max_freq, 1,
                                   tab.data_ptr(), st) == 0
    assert torch.equal(zs[..., :zw], z)
    torch.testing.assert_close((zs[..., :C].float() + zs[..., zw:zw + C].float()).cpu(), want, rtol=2e-5, atol=2e-5)
    if zw == 32 and 17 <= C <= 23:
        # merged tail (xattn_small.cu): the lo half also carries the hi parts of columns 16..C-1 right behind its zero
        # column C; everything else above the context width is zero
        e = C - 16
        assert bool((zs[..., zw + C] == 0).all())
        assert torch.equal(zs[..., zw + C + 1:zw + C + 1 + e], zs[..., 16:C])
        assert bool((zs[..., zw + C + 1 + e:] == 0).all())
    else:
        assert bool((zs[..., zw + C:] == 0).all())
    ldz = (C + 7) // 8 * 8
    z2 = torch.full((b * N, ldz), float("nan"), dtype=torch.float16, device="cuda")
    assert lib.hn_op_build_context(raw.data_ptr(), z2.data_ptr(), ldz, 0, b, c, len(axes), sizes, bands, max_freq, 1,
                                   tab.data_ptr(), st) == 0
    torch.testing.assert_close(z2[:, :C].float().cpu().reshape(b, N, C), want, rtol=2e-3, atol=2e-3)


def _attention_ref(q, k, v, mask=None):
    """q (b,H,L,d) pre-scaled in log2 units, k/v (b,H,N,d) -> softmax over N in base 2."""
    s = q.double() @ k.double().transpose(-1, -2)
    if mask is not None:
        s = s.masked_fill(~mask[:, None, None, :], float("-inf"))
    p = torch.softmax(s * math.log(2.0), dim=-1)
    return (p @ v.double()).float()


@pytest.mark.parametrize("b,H,L,N,nsplit", [(1, 1, 128, 64, 1), (2, 3, 130, 1000, 3), (1, 2, 25, 4097, 5),
                                            (2, 8, 256, 20000, 0)])
@pytest.mark.parametrize("masked", [False, True])
def test_generic_attention_kernel(b, H, L, N, nsplit, masked):
    lib, st = _lib_stream()
    g = torch.Generator(device="cuda").manual_seed(N)
    q = torch.randn(b, L, H * 64, device="cuda", generator=g).half()
    kv = torch.randn(b, N, 2 * H * 64, device="cuda", generator=g).half()
    kv[..., :H * 64] *= 0.3
    mask = (torch.rand(b, N, device="cuda", generator=g) > 0.3) if masked else None
    if masked:
        mask[:, :3] = True
    if nsplit == 0:
        nsplit = lib.hn_op_attention_nsplit(b, L, H, N, 0)
    n_lt = (L + 127) // 128
    acc = torch.full((b, nsplit, H, n_lt * 128, 64), float("nan"), device="cuda")
    ml = torch.full((b, nsplit, H, n_lt * 128, 2), float("nan"), device="cuda")
    acc_v = acc.view(-1)[: b * nsplit * H * L * 64]
    ml_v = ml.view(-1)[: b * nsplit * H * L * 2]
    bits = torch.zeros(b * ((N + 63) // 64), dtype=torch.int64, device="cuda")
    mk = mask.to(torch.uint8).contiguous() if masked else None
    rc = lib.hn_op_attention(q.data_ptr(), H * 64, kv.data_ptr(), 2 * H * 64, 0, H * 64, 0, 0, 64, b, L, H, N, nsplit,
                             mk.data_ptr() if masked else None, bits.data_ptr(), acc_v.data_ptr(), ml_v.data_ptr(), st)
    assert rc == 0, _lib.last_error()
    out = torch.zeros(b * L, H * 64, dtype=torch.float16, device="cuda")
    assert lib.hn_op_combine(acc_v.data_ptr(), ml_v.data_ptr(), b, nsplit, H, L, 0, 0, 64, 64, None, None, out.data_ptr(),
                             H * 64, st) == 0
    qh = q.float().view(b, L, H, 64).permute(0, 2, 1, 3)
    kh = kv[..., :H * 64].float().view(b, N, H, 64).permute(0, 2, 1, 3)
    vh = kv[..., H * 64:].float().view(b, N, H, 64).permute(0, 2, 1, 3)
    want = _attention_ref(qh, kh, vh, mask).permute(0, 2, 1, 3).reshape(b * L, H * 64)
    torch.testing.assert_close(out.float(), want, rtol=3e-3, atol=3e-3)  # fp16 P and fp16 output


def _split_cols(x):
    """(..., w) fp32 -> (..., 2w) fp16 [hi | lo]."""
    hi = x.half()
    return torch.cat([hi, (x - hi.float()).half()], dim=-1)


@pytest.mark.parametrize("kd,C", [(32, 18), (32, 31), (64, 50), (32, 13), (32, 16), (32, 23), (64, 40)])
@pytest.mark.parametrize("b,H,L,N", [(1, 1, 128, 64), (2, 8, 512, 30000), (1, 3, 200, 4100), (2, 2, 130, 777)])
@pytest.mark.parametrize("masked", [False, True])
def test_small_context_attention_kernel(kd, C, b, H, L, N, masked):
    """xattn_small.cu: Q' (b, L, H*kd) with column C zero, z (b, N, kd) with column C one; accumulator column C
    must come back as the softmax denominator. Variant 1: single fp16 operands; variant 3 (the forward's mode):
    split operands, whose scores must match an fp64 product of the UNROUNDED inputs much more tightly."""
    lib, st = _lib_stream()
    g = torch.Generator(device="cuda").manual_seed(N + C)
    q32 = torch.zeros(b, L, H, kd, device="cuda")
    q32[..., :C] = torch.randn(b, L, H, C, device="cuda", generator=g) * 0.7
    z32 = torch.zeros(b, N, kd, device="cuda")
    z32[..., :C] = torch.randn(b, N, C, device="cuda", generator=g)
    z32[..., C] = 1.0
    mask = None
    if masked:
        mask = torch.rand(b, N, device="cuda", generator=g) > 0.4
        mask[:, :2] = True
        if N > 200:
            mask[0, 64:192] = False  # whole tiles masked out
    # variant 4 (the forward's mode for 17 <= C <= 23): variant 3 on z rows that carry the merged tail, i.e. the hi
    # parts of columns 16..C-1 copied behind column C of the lo half — five score UMMAs per tile instead of six
    for variant in (1, 3, 4) if (kd == 32 and 17 <= C <= 23) else (1, 3):
        if variant == 1:
            q = q32.reshape(b, L, H * kd).half()
            z = z32.half()
            q_ld, kv_ld = H * kd, kd
            q_ref, z_ref = q.float().view(b, L, H, kd), z.float()
        else:
            qs = _split_cols(q32)                                          # (b, L, H, 2kd)
            q = torch.cat([qs[..., :kd].reshape(b, L, H * kd), qs[..., kd:].reshape(b, L, H * kd)], dim=-1).contiguous()
            z = _split_cols(z32).contiguous()
            if variant == 4:
                z[..., kd + C + 1:kd + C + 1 + (C - 16)] = z[..., 16:C]
            q_ld, kv_ld = 2 * H * kd, 2 * kd
            q_ref, z_ref = q32, z32
        qh = q_ref.permute(0, 2, 1, 3)
        zz = z_ref[:, None].expand(b, H, N, kd)
        # values: the kernel contracts P with the hi part of z
        zv = z32.half().float()[:, None].expand(b, H, N, kd)
        s = qh.double() @ zz.double().transpose(-1, -2)
        if mask is not None:
            s = s.masked_fill(~mask[:, None, None, :], float("-inf"))
        want = (torch.softmax(s * math.log(2.0), dim=-1) @ zv.double()).float()  # columns < C = sum p z, column C = 1
        nsplit = lib.hn_op_attention_nsplit(b, L, H, N, kd)
        n_lt = (L + 127) // 128
        acc = torch.full((b * nsplit * H * n_lt * 128 * kd,), float("nan"), device="cuda")
        ml = torch.full((b * nsplit * H * n_lt * 128 * 2,), float("nan"), device="cuda")
        bits = torch.zeros(b * ((N + 63) // 64), dtype=torch.int64, device="cuda")
        mk = mask.to(torch.uint8).contiguous() if masked else None
        rc = lib.hn_op_attention(q.data_ptr(), q_ld, z.data_ptr(), kv_ld, 0, 0, variant, C, 64, b, L, H, N, nsplit,
                                 mk.data_ptr() if masked else None, bits.data_ptr(), acc.data_ptr(), ml.data_ptr(), st)
        assert rc == 0, _lib.last_error()
        # combine with an identity V projection: Wv = I (C x C), bias 0 -> O[:, h*64 + d] = u_d / den for d < C
        dh = min(C, 64)
        Wv = torch.zeros(H * dh, kd, device="cuda")
        for h in range(H):
            Wv[h * dh:(h + 1) * dh, :dh] = torch.eye(dh, device="cuda")
        bv = torch.zeros(H * dh, device="cuda")
        out = torch.zeros(b * L, H * 64, dtype=torch.float16, device="cuda")
        assert lib.hn_op_combine(acc.data_ptr(), ml.data_ptr(), b, nsplit, H, L, C, kd, dh, 64, Wv.data_ptr(), bv.data_ptr(),
                                 out.data_ptr(), H * 64, st) == 0
        got = out.float().view(b, L, H, 64).permute(0, 2, 1, 3)[..., :dh]
        torch.testing.assert_close(got, want[..., :dh], rtol=3e-3, atol=3e-3)


def test_small_context_split_scores_are_exact_for_peaked_softmax():
    """Large-magnitude scores (|s| up to ~60 log2 units) over a long axis: with single fp16 operands the softmax
    weights are off by |s| 2^-11; the split operands (variant 3) must reproduce the fp64 softmax of the unrounded
    inputs an order of magnitude more tightly."""
    lib, st = _lib_stream()
    b, H, L, N, kd, C = 1, 2, 128, 40000, 32, 18
    g = torch.Generator(device="cuda").manual_seed(7)
    q32 = torch.zeros(b, L, H, kd, device="cuda")
    q32[..., :C] = torch.randn(b, L, H, C, device="cuda", generator=g) * 3.5
    z32 = torch.zeros(b, N, kd, device="cuda")
    z32[..., :C] = torch.randn(b, N, C, device="cuda", generator=g)
    z32[..., C] = 1.0
    s = q32.permute(0, 2, 1, 3).double() @ z32[:, None].expand(b, H, N, kd).double().transpose(-1, -2)
    p = torch.softmax(s * math.log(2.0), dim=-1)
    want = (p @ z32.half().float()[:, None].expand(b, H, N, kd).double()).float()
    errs = {}
    for variant in (1, 3, 4):
        if variant == 1:
            q, z, q_ld, kv_ld = q32.reshape(b, L, H * kd).half(), z32.half(), H * kd, kd
        else:
            qs = _split_cols(q32)
            q = torch.cat([qs[..., :kd].reshape(b, L, H * kd), qs[..., kd:].reshape(b, L, H * kd)], dim=-1).contiguous()
            z, q_ld, kv_ld = _split_cols(z32).contiguous(), 2 * H * kd, 2 * kd
            if variant == 4:   # merged tail (the forward's mode at this context width)
                z[..., kd + C + 1:kd + C + 1 + (C - 16)] = z[..., 16:C]
        nsplit = lib.hn_op_attention_nsplit(b, L, H, N, kd)
        acc = torch.full((b * nsplit * H * 128 * kd,), float("nan"), device="cuda")
        ml = torch.full((b * nsplit * H * 128 * 2,), float("nan"), device="cuda")
        assert lib.hn_op_attention(q.data_ptr(), q_ld, z.data_ptr(), kv_ld, 0, 0, variant, C, 64, b, L, H, N, nsplit, None,
                                   None, acc.data_ptr(), ml.data_ptr(), st) == 0, _lib.last_error()
        Wv = torch.zeros(H * C, kd, device="cuda")
        for h in range(H):
            Wv[h * C:(h + 1) * C, :C] = torch.eye(C, device="cuda")
        out = torch.zeros(b * L, 2 * H * 64, dtype=torch.float16, device="cuda")
        assert lib.hn_op_combine(acc.data_ptr(), ml.data_ptr(), b, nsplit, H, L, C, kd, C, 64, Wv.data_ptr(),
                                 torch.zeros(H * C, device="cuda").data_ptr(), out.data_ptr(), 2 * H * 64, st) == 0
        got = out[:, :H * 64].float().view(b, L, H, 64).permute(0, 2, 1, 3)[..., :C]
        errs[variant] = float((got - want[..., :C]).abs().max())
    print("peaked softmax, max abs error of sum p z: single fp16", errs[1], "split", errs[3], "split, merged tail", errs[4])
    assert errs[3] < 2.5e-3 and errs[4] < 2.5e-3            # fp16 P and fp16 output remain
    assert errs[3] < 0.5 * errs[1] or errs[1] < 2.5e-3
    assert errs[4] < 0.5 * errs[1] or errs[1] < 2.5e-3


@pytest.mark.parametrize("variant", [1, 3])
def test_small_context_attention_raises_reference_max(variant):
    """Scores that keep growing along the token axis force the lazy reference max to be raised many times
    (exact path, accumulator rescale, re-folded offset in Q') — result must still be the exact softmax."""
    lib, st = _lib_stream()
    b, H, L, N, kd, C = 1, 2, 128, 64 * 40, 32, 4
    q = torch.zeros(b, L, H, kd, device="cuda")
    q[..., 0] = torch.linspace(0.5, 2.0, L, device="cuda")[None, :, None]
    q = q.reshape(b, L, H * kd).half()
    z = torch.zeros(b, N, kd, device="cuda")
    z[..., 0] = torch.linspace(-40, 60, N, device="cuda")      # scores sweep ~ -80 .. +120 in log2 units
    z[..., 1] = torch.randn(b, N, device="cuda")
    z[..., C] = 1.0
    z = z.half()
    qh = q.float().view(b, L, H, kd).permute(0, 2, 1, 3)
    zz = z.float()[:, None].expand(b, H, N, kd)
    want = _attention_ref(qh, zz, zz)
    nsplit = 2
    acc = torch.full((b * nsplit * H * 128 * kd,), float("nan"), device="cuda")
    ml = torch.full((b * nsplit * H * 128 * 2,), float("nan"), device="cuda")
    q_ld, kv_ld = H * kd, kd
    if variant == 3:  # split layout of the same (already fp16-exact) operands: lo parts are zero
        q = torch.cat([q, torch.zeros_like(q)], dim=-1).contiguous()
        z = torch.cat([z, torch.zeros_like(z)], dim=-1).contiguous()
        q_ld, kv_ld = 2 * H * kd, 2 * kd
    assert lib.hn_op_attention(q.data_ptr(), q_ld, z.data_ptr(), kv_ld, 0, 0, variant, C, 64, b, L, H, N, nsplit, None,
                               None, acc.data_ptr(), ml.data_ptr(), st) == 0
    Wv = torch.zeros(H * C, kd, device="cuda")
    for h in range(H):
        Wv[h * C:(h + 1) * C, :C] = torch.eye(C, device="cuda")
    out = torch.zeros(b * L, H * 64, dtype=torch.float16, device="cuda")
    assert lib.hn_op_combine(acc.data_ptr(), ml.data_ptr(), b, nsplit, H, L, C, kd, C, 64, Wv.data_ptr(),
                             torch.zeros(H * C, device="cuda").data_ptr(), out.data_ptr(), H * 64, st) == 0
    got = out.float().view(b, L, H, 64).permute(0, 2, 1, 3)[..., :C]
    assert bool(torch.isfinite(got).all())
    torch.testing.assert_close(got, want[..., :C], rtol=5e-3, atol=5e-3)
