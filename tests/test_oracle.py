"""CPU: pins the oracle (oracle/healnet_oracle.py) against fixtures produced by executing the unmodified
reference (tests/golden/make_golden.py). The reference's own tests hold no numeric vectors for this path
(healnet/tests/test_healnet.py:26-67 asserts shapes only), so these reference-run outputs are the pin."""
import pytest
import torch

from oracle import healnet_oracle as O

CASES = ["tri_small", "omic_wsi_tied", "plain_no_head", "two_ltiles", "prod_ucec", "wide_heads"]


def _cfg(kwargs):
    kw = {k: v for k, v in kwargs.items() if k in O.OracleConfig.__dataclass_fields__}
    return O.OracleConfig(**kw)


def _inputs(ins, n):
    return [ins[str(i)] for i in range(n)]


@pytest.mark.parametrize("name", CASES)
def test_forward_matches_reference(golden, name):
    meta, sd, ins, outs, _ = golden(name)
    cfg = _cfg(meta["kwargs"])
    x = _inputs(ins, cfg.n_modalities)
    lat = O.forward(sd, cfg, x, return_embeddings=True)
    torch.testing.assert_close(lat, outs["latents"], rtol=1e-4, atol=2e-5)
    if "logits" in outs:
        torch.testing.assert_close(O.forward(sd, cfg, x), outs["logits"], rtol=1e-4, atol=2e-5)
    # head-chunked evaluation (used by the CPU baseline to bound memory) is the same arithmetic
    torch.testing.assert_close(O.forward(sd, cfg, x, return_embeddings=True, head_chunk=1), lat, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", CASES)
def test_missing_modality_semantics(golden, name):
    """healnet.py:229-245: default skips only the cross block of a missing modality; verbose=True skips the
    latent block too; a short list behaves like trailing Nones."""
    meta, sd, ins, outs, _ = golden(name)
    cfg = _cfg(meta["kwargs"])
    x = _inputs(ins, cfg.n_modalities)
    miss = [x[0], None] + x[2:]
    torch.testing.assert_close(O.forward(sd, cfg, miss, return_embeddings=True), outs["missing1_latents"],
                               rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(O.forward(sd, cfg, miss, return_embeddings=True, verbose=True),
                               outs["missing1_verbose_latents"], rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(O.forward(sd, cfg, [x[0]], return_embeddings=True), outs["short_list_latents"],
                               rtol=1e-4, atol=2e-5)
    if cfg.self_per_cross_attn:
        assert not torch.allclose(outs["missing1_latents"], outs["missing1_verbose_latents"])


def test_mask(golden):
    meta, sd, ins, outs, _ = golden("masked")
    cfg = _cfg(meta["kwargs"])
    out = O.forward(sd, cfg, [ins["0"]], mask=ins["mask"])
    torch.testing.assert_close(out, outs["logits"], rtol=1e-4, atol=2e-5)
    assert not torch.allclose(O.forward(sd, cfg, [ins["0"]]), outs["logits"], atol=1e-4)


def test_attention_module(golden):
    meta, sd, ins, outs, extra = golden("attention")
    kw = meta["kwargs"]
    y, w = O.attention(ins["x"], ins["context"], sd["to_q.weight"], sd["to_kv.weight"], sd["to_out.0.weight"],
                       sd["to_out.0.bias"], kw["heads"], want_weights=True)
    torch.testing.assert_close(y, outs["cross"], rtol=1e-4, atol=2e-5)
    # last call of the fixture was the masked one, whose attention matrix it kept (healnet.py:420)
    ym, wm = O.attention(ins["x"], ins["context"], sd["to_q.weight"], sd["to_kv.weight"], sd["to_out.0.weight"],
                         sd["to_out.0.bias"], kw["heads"], mask=ins["mask"], want_weights=True)
    torch.testing.assert_close(ym, outs["cross_masked"], rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(wm, extra["attn/cross"], rtol=1e-4, atol=1e-6)
    sd2 = {k.split("/", 1)[1]: v for k, v in extra.items() if k.startswith("sd2/")}
    ys = O.attention(ins["x"], ins["x"], sd2["to_q.weight"], sd2["to_kv.weight"], sd2["to_out.0.weight"],
                     sd2["to_out.0.bias"], meta["kwargs_self"]["heads"])
    torch.testing.assert_close(ys, outs["self"], rtol=1e-4, atol=2e-5)


def test_softmax_temperature_is_two_over_sqrt_d(golden):
    """healnet.py:375,409,419: effective logit scale is 2/sqrt(dim_head)."""
    torch.manual_seed(0)
    q, k = torch.randn(1, 5, 16), torch.randn(1, 7, 16)
    eye = torch.eye(16)
    y, w = O.attention(q, k, eye, torch.cat([eye, eye]), eye, torch.zeros(16), heads=1, want_weights=True)
    torch.testing.assert_close(w[0], torch.softmax(2.0 / 4.0 * q[0] @ k[0].t(), dim=-1))


def test_fourier_tables(golden):
    _, _, _, _, extra = golden_fourier()
    for key, ref in extra.items():
        size = int(key.split("/")[1])
        torch.testing.assert_close(O.fourier_table(size, 10.0, 2), ref, rtol=0, atol=1e-6)
        assert ref.shape == (size, 5)
    assert float(extra["fourier/1"][0, -1]) == -1.0  # a size-1 axis sits at position -1 (linspace(-1,1,1))


def golden_fourier():
    import numpy as np, os
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, "fourier.npz"))
    return None, None, None, None, {k: torch.from_numpy(z[k]) for k in z.files}


def test_flop_model_matches_survey():
    """SURVEY.md section 8d values: cfg 1 = 2.200 TFLOP/sample, README-default latent = 0.586."""
    shapes = [(1,), (224, 224), (12, 224, 224)]
    cfg = O.OracleConfig(n_modalities=3, channel_dims=[2000, 3, 3], num_spatial_axes=[1, 2, 3], out_dims=4,
                         l_c=512, l_d=512)
    assert abs(O.flops_per_sample(cfg, shapes) / 1e12 - 2.200) < 2e-3
    cfg.l_c = cfg.l_d = 128
    assert abs(O.flops_per_sample(cfg, shapes) / 1e12 - 0.586) < 2e-3


def test_bench_flop_model_agrees_with_the_oracle_and_survey():
    """bench.py's GPU arm carries its own work model (it does not import the oracle): same numbers for every
    BASELINE configuration, and the SURVEY.md section 8d values (cfg 2 0.0573, cfg 3 6.369, cfg 4 0.1162, cfg 5 0.4401)."""
    import bench
    want = {"cfg1": 2.200, "cfg2": 0.0573, "cfg3": 6.369, "cfg4": 0.1162, "cfg5": 0.4401}
    for name, (kw, shapes, _) in bench.WORKLOADS.items():
        cfg = O.OracleConfig(**{k: v for k, v in kw.items() if k in O.OracleConfig.__dataclass_fields__})
        got = bench.flops_per_sample(kw, shapes)
        assert abs(got - O.flops_per_sample(cfg, [s[:-1] for s in shapes])) < 1.0
        if name in want:
            assert abs(got / 1e12 - want[name]) < 2e-3 * max(1.0, want[name]), name
