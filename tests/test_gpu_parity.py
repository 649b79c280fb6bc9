"""GPU (B200): parity of the CUDA path, called through the drop-in module -> C ABI, against
  (a) the committed reference-run fixtures (tests/golden), and
  (b) the oracle on seeded inputs at sizes it finishes in seconds.
Tolerance (BASELINE.json north_star): rtol 1e-3 / atol 1e-4 in fp32 for logits; the latent array is
compared at the same rtol with atol 5e-4 (it is 2-3 orders of magnitude larger in scale than the logits'
atol: |latents| ~ 1..10 after `depth` residual updates)."""
import copy

import pytest
import torch

import healnet_b200
from healnet_b200 import Attention, HealNet
from oracle import healnet_oracle as O

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-3, 1e-4
LAT_ATOL = 5e-4
CASES = ["tri_small", "omic_wsi_tied", "plain_no_head", "two_ltiles", "prod_ucec", "wide_heads"]


def _model(meta, sd):
    m = HealNet(**meta["kwargs"])
    m.load_state_dict(sd)
    return m.cuda().eval()


def _inputs(ins, n):
    return [ins[str(i)].cuda() for i in range(n)]


def _cfg(kwargs):
    return O.OracleConfig(**{k: v for k, v in kwargs.items() if k in O.OracleConfig.__dataclass_fields__})


@pytest.mark.parametrize("name", CASES)
def test_golden_forward(golden, name):
    meta, sd, ins, outs, _ = golden(name)
    m = _model(meta, sd)
    x = _inputs(ins, meta["kwargs"]["n_modalities"])
    keep = [t.clone() for t in x]
    lat = m(list(x), return_embeddings=True)
    assert lat.device.type == "cuda" and lat.dtype == torch.float32
    torch.testing.assert_close(lat.cpu(), outs["latents"], rtol=RTOL, atol=LAT_ATOL)
    if "logits" in outs:
        torch.testing.assert_close(m(list(x)).cpu(), outs["logits"], rtol=RTOL, atol=ATOL)
    else:  # final_classifier_head=False -> Identity head returns the latent array (healnet.py:185)
        torch.testing.assert_close(m(list(x)).cpu(), outs["latents"], rtol=RTOL, atol=LAT_ATOL)
    for a, b in zip(x, keep):  # unlike the reference (healnet.py:222) the caller's tensors are untouched
        assert torch.equal(a, b)
    assert m.last_launch_count > 0


@pytest.mark.parametrize("name", CASES)
def test_golden_missing_modalities(golden, name):
    meta, sd, ins, outs, _ = golden(name)
    m = _model(meta, sd)
    x = _inputs(ins, meta["kwargs"]["n_modalities"])
    miss = [x[0], None] + x[2:]
    torch.testing.assert_close(m(miss, return_embeddings=True).cpu(), outs["missing1_latents"], rtol=RTOL, atol=LAT_ATOL)
    torch.testing.assert_close(m(miss, return_embeddings=True, verbose=True).cpu(), outs["missing1_verbose_latents"],
                               rtol=RTOL, atol=LAT_ATOL)
    torch.testing.assert_close(m([x[0]], return_embeddings=True).cpu(), outs["short_list_latents"], rtol=RTOL,
                               atol=LAT_ATOL)


def test_golden_mask(golden):
    meta, sd, ins, outs, _ = golden("masked")
    m = _model(meta, sd)
    out = m([ins["0"].cuda()], mask=ins["mask"].cuda())
    torch.testing.assert_close(out.cpu(), outs["logits"], rtol=RTOL, atol=ATOL)
    lat = m([ins["0"].cuda()], mask=ins["mask"].cuda(), return_embeddings=True)
    torch.testing.assert_close(lat.cpu(), outs["latents"], rtol=RTOL, atol=LAT_ATOL)


def test_single_token_mask_broadcasts_over_every_modality():
    """A (b, 1) mask broadcasts over the token axis of every modality (healnet.py:411-415): True is a no-op, False
    masks all tokens of that sample -> NaN rows, as in the reference (oracle)."""
    kw = dict(n_modalities=2, channel_dims=[40, 3], num_spatial_axes=[1, 2], out_dims=3, l_c=32, l_d=64, depth=1)
    torch.manual_seed(5)
    model = HealNet(**kw).eval()
    xs = [torch.rand(3, 1, 40), torch.rand(3, 50, 60, 3)]
    mask = torch.tensor([[True], [False], [True]])
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    want = O.forward(sd, _cfg(kw), xs, mask=mask)
    got = model.cuda()([t.cuda() for t in xs], mask=mask.cuda()).cpu()
    assert bool(torch.isnan(want[1]).all()) and bool(torch.isnan(got[1]).all())
    torch.testing.assert_close(got[[0, 2]], want[[0, 2]], rtol=RTOL, atol=ATOL)


def test_golden_attention_module(golden):
    meta, sd, ins, outs, extra = golden("attention")
    att = Attention(**meta["kwargs"])
    att.load_state_dict(sd)
    att = att.cuda().eval()
    x, ctx, mask = ins["x"].cuda(), ins["context"].cuda(), ins["mask"].cuda()
    torch.testing.assert_close(att(x, context=ctx).cpu(), outs["cross"], rtol=RTOL, atol=2e-4)
    torch.testing.assert_close(att(x, context=ctx, mask=mask).cpu(), outs["cross_masked"], rtol=RTOL, atol=2e-4)
    att2 = Attention(**meta["kwargs_self"])
    att2.load_state_dict({k.split("/", 1)[1]: v for k, v in extra.items() if k.startswith("sd2/")})
    torch.testing.assert_close(att2.cuda()(x).cpu(), outs["self"], rtol=RTOL, atol=2e-4)


def test_reference_unit_tests_shapes():
    """The bodies of the reference's healnet/tests/test_healnet.py:26-67 (CPU tensors, CPU-constructed
    modules) against the drop-in classes."""
    b, t_d, i_c, l_c, l_d = 10, 2189, 100, 256, 32
    query = torch.randn(b, 1, t_d)
    latent = torch.randn(b, l_c, l_d)
    attention = Attention(query_dim=l_d, context_dim=t_d)
    assert attention(x=latent, context=query).shape == (b, l_c, l_d)
    tabular = torch.randn(b, 1, t_d)
    image = torch.randn(b, 224, 224, i_c)
    m1 = HealNet(n_modalities=1, channel_dims=[t_d], num_spatial_axes=[1], out_dims=5)
    out1 = m1([tabular])
    assert out1.shape == (b, 5) and out1.device.type == "cpu" and bool(torch.isfinite(out1).all())
    m2 = HealNet(n_modalities=2, channel_dims=[t_d, i_c], num_spatial_axes=[1, 2], out_dims=4)
    out2 = m2([tabular, image])
    assert out2.shape == (b, 4) and bool(torch.isfinite(out2).all())
    with pytest.raises(AssertionError):
        HealNet(n_modalities=1, channel_dims=[t_d, i_c], num_spatial_axes=[1, 1], out_dims=4)


def _randomise(model, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("norm.weight") or n.endswith("norm_context.weight") or n == "to_logits.1.weight":
                p.copy_(1.0 + 0.3 * torch.randn(p.shape, generator=g))
            elif n.endswith(".bias"):
                p.copy_(0.3 * torch.randn(p.shape, generator=g))


ORACLE_CASES = {
    # README-shaped 3-modality model at reduced spatial extent (cfg 1 family): tab wide-context, img/vol small-C
    "readme_reduced": (dict(n_modalities=3, channel_dims=[2000, 3, 3], num_spatial_axes=[1, 2, 3], out_dims=4,
                            l_c=256, l_d=128), [(2, 1, 2000), (2, 56, 56, 3), (2, 4, 40, 40, 3)]),
    # cfg 2 family: omic + WSI patch features (generic K/V projection path), latent 256 x 512 at reduced N
    "omic_wsi": (dict(n_modalities=2, channel_dims=[2000, 1024], num_spatial_axes=[1, 1], out_dims=4, l_c=256,
                      l_d=512, depth=2), [(2, 1, 2000), (2, 700, 1024)]),
    # cfg 3 family: wide latents (l_d = 1024), depth 8, all three modalities at reduced spatial extent
    "wide_latents_deep": (dict(n_modalities=3, channel_dims=[500, 3, 3], num_spatial_axes=[1, 2, 3], out_dims=4,
                               l_c=128, l_d=1024, depth=8), [(2, 1, 500), (2, 48, 48, 3), (2, 3, 32, 32, 3)]),
    # tuned production hyper-parameters (config/best_hyperparams.yml:8-18,30): tiny odd latents, 1 head of 63
    # long token axis with a wide head: generic streaming path, two atoms per head, non-precise
    "wide_head_long_axis": (dict(n_modalities=1, channel_dims=[96], num_spatial_axes=[1], out_dims=4, l_c=64, l_d=96,
                                 x_heads=2, cross_dim_head=103, l_heads=2, latent_dim_head=72, depth=2),
                            [(2, 3000, 96)]),
    "production_odd": (dict(n_modalities=2, channel_dims=[300, 96], num_spatial_axes=[1, 1], out_dims=4, l_c=25,
                            l_d=119, x_heads=1, cross_dim_head=63, self_per_cross_attn=0),
                       [(3, 1, 300), (3, 333, 96)]),
}


@pytest.mark.parametrize("name", list(ORACLE_CASES))
def test_against_oracle(name):
    kw, shapes = ORACLE_CASES[name]
    torch.manual_seed(11)
    model = HealNet(**kw).eval()
    _randomise(model, 5)
    g = torch.Generator().manual_seed(3)
    xs = [torch.rand(s, generator=g) for s in shapes]
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    want = O.forward(sd, _cfg(kw), xs)
    want_lat = O.forward(sd, _cfg(kw), xs, return_embeddings=True)
    model.cuda()
    got = model([t.cuda() for t in xs]).cpu()
    got_lat = model([t.cuda() for t in xs], return_embeddings=True).cpu()
    torch.testing.assert_close(got_lat, want_lat, rtol=RTOL, atol=LAT_ATOL)
    torch.testing.assert_close(got, want, rtol=RTOL, atol=ATOL)


def test_repack_after_inplace_update_and_device_moves():
    kw, shapes = ORACLE_CASES["production_odd"]
    torch.manual_seed(2)
    model = HealNet(**kw).eval().cuda()
    xs = [torch.rand(s).cuda() for s in shapes]
    a = model(xs)
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(1.05)
    b = model(xs)
    assert not torch.allclose(a, b)
    sd = {k: v.cpu() for k, v in model.state_dict().items()}
    want = O.forward(sd, _cfg(kw), [t.cpu() for t in xs])
    torch.testing.assert_close(b.cpu(), want, rtol=RTOL, atol=ATOL)
    # CPU-resident module + CPU inputs: staged to the GPU, result returned on the CPU
    c = copy.deepcopy(model).cpu()([t.cpu() for t in xs])
    assert c.device.type == "cpu"
    torch.testing.assert_close(c, b.cpu(), rtol=1e-5, atol=1e-6)


def test_batch_independence_and_determinism():
    kw, shapes = ORACLE_CASES["readme_reduced"]
    torch.manual_seed(4)
    model = HealNet(**kw).eval().cuda()
    xs = [torch.rand(s).cuda() for s in shapes]
    full = model(xs)
    assert torch.equal(full, model(xs)), "forward must be deterministic (fixed split-combine order)"
    for i in range(xs[0].shape[0]):
        one = model([t[i:i + 1] for t in xs])
        torch.testing.assert_close(one, full[i:i + 1], rtol=1e-4, atol=2e-5)


def test_bf16_module_and_inputs():
    """BASELINE config 3 runs the model in bf16 (`model.bfloat16()`, bf16 inputs). The kernels keep fp32 master
    arithmetic (bf16 parameters / inputs are widened, the result is returned in bf16), so the output must sit
    within bf16 rounding of the fp32 oracle evaluated on the same bf16-rounded weights and inputs — stated bf16
    tolerance: atol 2e-2 (the reference's own bf16 run differs from its fp32 run by 8.5e-3)."""
    kw, shapes = ORACLE_CASES["readme_reduced"]
    torch.manual_seed(21)
    model = HealNet(**kw).eval().bfloat16()
    xs = [torch.rand(s).bfloat16() for s in shapes]
    sd = {k: v.float() for k, v in model.state_dict().items()}
    want = O.forward(sd, _cfg(kw), [t.float() for t in xs])
    got = model.cuda()([t.cuda() for t in xs])
    assert got.dtype == torch.bfloat16
    torch.testing.assert_close(got.float().cpu(), want, rtol=1e-2, atol=2e-2)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_native_16bit_inputs_equal_widened_inputs(dtype):
    """Inputs that all arrive in bf16 / fp16 are read as they are (hn_set_io_dtype: widened while the context rows are
    standardised). Widening is exact, so the result must be bit-identical to the forward of the same values passed
    as fp32 — on the small-context streaming path, the generic path and the tabular row."""
    kw, shapes = ORACLE_CASES["readme_reduced"]
    torch.manual_seed(13)
    model = HealNet(**kw).eval().cuda()
    xs = [torch.rand(s).to(dtype).cuda() for s in shapes]
    with torch.no_grad():
        a = model(list(xs))
        b = model([t.float() for t in xs])
        assert a.dtype == dtype and torch.equal(a.float(), b.to(dtype).float())
        a_lat = model(list(xs), return_embeddings=True)
        b_lat = model([t.float() for t in xs], return_embeddings=True)
        assert torch.equal(a_lat.float(), b_lat.to(dtype).float())
        kw2, shapes2 = ORACLE_CASES["omic_wsi"]
        m2 = HealNet(**kw2).eval().cuda()
        ys = [torch.rand(s).to(dtype).cuda() for s in shapes2]
        assert torch.equal(m2(list(ys)).float(), m2([t.float() for t in ys]).to(dtype).float())


def test_pinned_host_inputs_overlap_path():
    """Pinned host tensors take the side-stream copy + per-modality ready-event path (hn_forward_ex); the result must
    be bit-identical to the device-resident call, call after call."""
    kw, shapes = ORACLE_CASES["readme_reduced"]
    torch.manual_seed(8)
    model = HealNet(**kw).eval().cuda()
    host = [torch.rand(s).pin_memory() for s in shapes]
    want = model([t.cuda() for t in host])
    for _ in range(3):
        got = model(list(host))
        assert got.device.type == "cpu"
        assert torch.equal(got, want.cpu())
    miss = model([host[0], None, host[2]])
    assert torch.equal(miss, model([host[0].cuda(), None, host[2].cuda()]).cpu())


@pytest.mark.parametrize("name", ["tri_small", "wide_heads", "omic_wsi_tied"])
def test_attention_weight_export(golden, name):
    """Opt-in replacement of the reference's retained `attn_weights` (healnet.py:420, :252-262; consumed by
    explainer.py:102-104): one (b*h, L, N) matrix per Attention module, in module order, last call wins."""
    meta, sd, ins, outs, _ = golden(name)
    kw = meta["kwargs"]
    m = _model(meta, sd)
    x = _inputs(ins, kw["n_modalities"])
    assert all(w is None for w in m.get_attention_weights())
    m.export_attention_weights = True
    logits = m(list(x))
    got = m.get_attention_weights()
    # oracle: attention matrices in call order -> keep the last call of every module (tying, repeated self-attn)
    calls = []
    O.forward(sd, _cfg(kw), [t.cpu() for t in x], collect_weights=calls)
    M, depth, spc = kw["n_modalities"], kw.get("depth", 3), kw.get("self_per_cross_attn", 1)
    per_module = {}
    it = iter(calls)
    for l in range(depth):
        for i in range(M):
            per_module[id(m.layers[l][2 * i].fn)] = next(it)
            if spc:
                per_module[id(m.layers[l][-1][0].fn)] = next(it)
    mods = [mod for mod in m.modules() if isinstance(mod, Attention)]
    assert len(got) == len(mods)
    for mod, w in zip(mods, got):
        want = per_module[id(mod)]
        assert w is not None and tuple(w.shape) == tuple(want.shape)
        torch.testing.assert_close(w.cpu(), want, rtol=5e-3, atol=1e-5)
        torch.testing.assert_close(w.sum(-1).cpu(), torch.ones(w.shape[:2]), rtol=1e-3, atol=1e-3)
    # export must not change the result, and switching it off stops refreshing the tensors
    m.export_attention_weights = False
    torch.testing.assert_close(m(list(x)), logits, rtol=0, atol=0)


def test_attention_weight_export_long_axis_and_mask():
    """Small-context streaming path (N > 2048) with a mask: exported rows are the applied softmax, zeros at masked tokens."""
    torch.manual_seed(3)
    kw = dict(n_modalities=1, channel_dims=[3], num_spatial_axes=[2], out_dims=2, depth=1, l_c=40, l_d=64, x_heads=2,
              cross_dim_head=32, self_per_cross_attn=0)
    m = HealNet(**kw).eval().cuda()
    x = torch.rand(2, 50, 60, 3, device="cuda")
    mask = torch.rand(2, 3000, device="cuda") > 0.3
    m.export_attention_weights = True
    m([x], mask=mask)
    (w,) = m.get_attention_weights()
    calls = []
    O.forward({k: v.cpu() for k, v in m.state_dict().items()}, _cfg(kw), [x.cpu()], mask=mask.cpu(), collect_weights=calls)
    assert tuple(w.shape) == (4, 40, 3000)
    torch.testing.assert_close(w.cpu(), calls[0], rtol=5e-3, atol=1e-6)
    assert bool((w[:2][:, :, ~mask[0]] == 0).all())
    m.export_attention_max_bytes = 1000
    with pytest.raises(MemoryError):
        m([x])


def test_standalone_attention_reuses_its_packed_weights_until_they_change():
    """Attention keeps the packed weights in its workspace (hn_attention_forward_cached): a second call with unchanged
    parameters skips the packing launches and must give the same result; an in-place parameter update (optimizer step)
    must be picked up."""
    from healnet_b200 import Attention
    torch.manual_seed(0)
    att = Attention(query_dim=96, context_dim=40, heads=4, dim_head=24).cuda()
    x, ctx = torch.randn(2, 50, 96, device="cuda"), torch.randn(2, 300, 40, device="cuda")
    with torch.no_grad():
        a = att(x, context=ctx)
        b = att(x, context=ctx)              # cached weights
        assert torch.equal(a, b)
        att.to_q.weight.mul_(1.5)            # in-place update bumps the parameter's version
        c = att(x, context=ctx)
        fresh = Attention(query_dim=96, context_dim=40, heads=4, dim_head=24).cuda()
        fresh.load_state_dict(att.state_dict())
        want = fresh(x, context=ctx)
    assert not torch.equal(a, c)
    assert torch.equal(c, want)
