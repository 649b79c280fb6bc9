"""GPU (B200): parity of the CUDA path against the oracle AT THE BENCHMARKED SIZES — every BASELINE.json config at
its full token counts, latent size and depth (one or two samples per case; the path is per-sample).

What the reduced-size cases of test_gpu_parity.py cannot show: the split-N regime of the streaming kernels (cfg 1
volume: 602 112 tokens = 9 408 tiles over ~37 splits), the lazily raised reference max over thousands of tiles, the
fp16 P / half2-polynomial error accumulating in the fp32 denominator, and large-magnitude (peaked) logits at scale.

Checkers:
  * `cfg1_full` — the oracle on the HOST CPU in fp32 (head_chunk=2: two heads of the 9.87 GB attention matrix at a
    time; ~10 s, ~8 GB), the authoritative comparison for the benchmarked workload. The same case also checks that the
    oracle evaluated on CUDA tensors in float64 agrees with the CPU run, which licenses its use below;
  * every other case — the SAME oracle code (oracle/healnet_oracle.py) on CUDA tensors in float64 (torch eager ops;
    the oracle is device- and dtype-agnostic), because the CPU run of cfg 3 / cfg 5 takes minutes.
Tolerance: BASELINE.json north star, rtol 1e-3 / atol 1e-4 on the logits (latent array: atol 5e-4, its scale is
1..10); cfg 3 (bf16 parameters and inputs): the stated bf16 tolerance rtol 1e-2 / atol 2e-2.
Every case appends its measured errors to gpurun_out/parity_fullsize.jsonl (summarised under profiles/)."""
import json
import os

import pytest
import torch

from healnet_b200 import HealNet
from oracle import healnet_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL, ATOL, LAT_ATOL = 1e-3, 1e-4, 5e-4

CFG = {
    "cfg1": (dict(n_modalities=3, channel_dims=[2000, 3, 3], num_spatial_axes=[1, 2, 3], out_dims=4, l_c=512, l_d=512),
             [(1, 2000), (224, 224, 3), (12, 224, 224, 3)]),
    "cfg2": (dict(n_modalities=2, channel_dims=[2000, 1024], num_spatial_axes=[1, 1], out_dims=4, l_c=256, l_d=512),
             [(1, 2000), (4096, 1024)]),
    "cfg3": (dict(n_modalities=3, channel_dims=[2000, 3, 3], num_spatial_axes=[1, 2, 3], out_dims=4, l_c=512,
                  l_d=1024, depth=8),
             [(1, 2000), (224, 224, 3), (12, 224, 224, 3)]),
    "cfg4": (dict(n_modalities=2, channel_dims=[2000, 768], num_spatial_axes=[1, 1], out_dims=4, l_c=512, l_d=512),
             [(1, 2000), (8192, 768)]),
    "cfg5": (dict(n_modalities=1, channel_dims=[512], num_spatial_axes=[1], out_dims=4, l_c=512, l_d=512),
             [(65536, 512)]),
}


def _cfg(kw):
    return O.OracleConfig(**{k: v for k, v in kw.items() if k in O.OracleConfig.__dataclass_fields__})


def _randomise(model, seed, q_gain=1.0):
    """LayerNorm affines and every bias re-drawn (default init leaves them at 1 / 0, which hides folding bugs);
    q_gain scales every cross-attention to_q so that the softmax becomes peaked (|logit| grows q_gain-fold)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("norm.weight") or n.endswith("norm_context.weight") or n == "to_logits.1.weight":
                p.copy_(1.0 + 0.3 * torch.randn(p.shape, generator=g))
            elif n.endswith(".bias"):
                p.copy_(0.3 * torch.randn(p.shape, generator=g))
            elif q_gain != 1.0 and n.endswith("fn.to_q.weight"):
                p.mul_(q_gain)


def _record(name, **kw):
    d = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_fullsize.jsonl"), "a") as f:
            f.write(json.dumps(dict(case=name, **kw)) + "\n")
    except OSError:
        pass
    print("parity", name, kw)


def _errs(got, want):
    d = (got.double() - want.double()).abs()
    return dict(max_abs=float(d.max()), max_rel=float((d / want.double().abs().clamp_min(1e-3)).max()),
                ref_absmax=float(want.abs().max()))


def _head(sd, lat):
    """to_logits of the oracle (healnet.py:181-185) applied to a latent array it returned."""
    g = lambda k: sd[k].to(lat.dtype)
    pooled = O.layer_norm(lat.mean(dim=1), g("to_logits.1.weight"), g("to_logits.1.bias"))
    return pooled @ g("to_logits.2.weight").t() + g("to_logits.2.bias")


def _oracle(sd, kw, xs, device, dtype):
    """oracle/healnet_oracle.py on `device` tensors in `dtype`, two heads at a time -> (logits, latents), fp32 CPU."""
    sdd = {k: v.to(device=device, dtype=dtype) for k, v in sd.items()}
    logits, lats = [], []
    with torch.no_grad():
        for i in range(xs[0].shape[0]):   # sample by sample (the path is per-sample): bounds the oracle's memory
            lat = O.forward(sdd, _cfg(kw), [t[i:i + 1].to(device=device, dtype=dtype) for t in xs], dtype=dtype,
                            head_chunk=2, return_embeddings=True)
            logits.append(_head(sdd, lat).float().cpu())
            lats.append(lat.float().cpu())
            del lat
    out = torch.cat(logits), torch.cat(lats)
    del sdd
    if device != "cpu":
        torch.cuda.empty_cache()
    return out


def _run(name, kw, shapes, batch, seed, q_gain=1.0, cpu_oracle=False, bf16=False, rtol=RTOL, atol=ATOL,
         lat_atol=LAT_ATOL):
    torch.manual_seed(seed)
    model = HealNet(**kw).eval()
    _randomise(model, seed + 1, q_gain)
    g = torch.Generator().manual_seed(seed + 2)
    xs = [torch.rand((batch,) + tuple(s), generator=g) for s in shapes]
    if bf16:
        model = model.bfloat16()
        xs = [t.bfloat16() for t in xs]
    sd = {k: v.detach().float().clone() for k, v in model.state_dict().items()}
    xs32 = [t.float() for t in xs]
    want64, want64_lat = _oracle(sd, kw, xs32, "cuda", torch.float64)
    rec = {}
    if cpu_oracle:
        want_cpu, want_cpu_lat = _oracle(sd, kw, xs32, "cpu", torch.float32)
        rec["cpu32_vs_cuda64"] = _errs(want_cpu, want64)
        # the two evaluations of the oracle must agree far inside the tolerance they are used to check
        torch.testing.assert_close(want_cpu, want64, rtol=1e-4, atol=2e-5)
        torch.testing.assert_close(want_cpu_lat, want64_lat, rtol=1e-4, atol=5e-5)
    model.cuda()
    dev_in = [t.cuda() for t in xs]
    got = model(list(dev_in)).float().cpu()
    got_lat = model(list(dev_in), return_embeddings=True).float().cpu()
    rec["logits"] = _errs(got, want64)
    rec["latents"] = _errs(got_lat, want64_lat)
    _record(name, batch=batch, q_gain=q_gain, launches=model.last_launch_count, **rec)
    if cpu_oracle:
        torch.testing.assert_close(got, want_cpu, rtol=rtol, atol=atol)
        torch.testing.assert_close(got_lat, want_cpu_lat, rtol=rtol, atol=lat_atol)
    torch.testing.assert_close(got_lat, want64_lat, rtol=rtol, atol=lat_atol)
    torch.testing.assert_close(got, want64, rtol=rtol, atol=atol)


def test_cfg1_full_sample_vs_cpu_oracle():
    """BASELINE cfg 1 (the benchmarked workload), one whole sample: tab 1x2000 + img 224x224x3 + vol 12x224x224x3,
    latent 512x512, depth 3 — CUDA path vs the oracle on the host CPU (and vs the float64 CUDA evaluation)."""
    kw, shapes = CFG["cfg1"]
    _run("cfg1_full", kw, shapes, batch=1, seed=100, cpu_oracle=True)


@pytest.mark.parametrize("gain", [8.0, 32.0])
def test_cfg1_full_sample_peaked_attention(gain):
    """Same shapes with every cross-attention `to_q` scaled: logits grow gain-fold, the softmax over the 602 112
    voxels becomes peaked, the lazily raised reference max has to be raised repeatedly inside and across splits, and
    any error of the fp16 score operands is no longer averaged away. x8 (|logit| up to ~26 nats) is held to the
    north-star tolerance. x32 (|logit| ~ 100 nats: a handful of voxels share a row's whole weight) is a stress
    characterisation: the scores are still fp32-exact (split operands), but the softmax weights are carried to the
    tensor cores as fp16 (relative rounding 5e-4, polynomial exp2 6e-4 rms), which a sum over two or three comparable
    voxels no longer averages away — measured latent error 1.8e-3 on values up to 15.7 (logits 1.3e-5); its bound is
    stated here as atol 3e-3 on the latent array, the logits keep the north-star tolerance."""
    kw, shapes = CFG["cfg1"]
    _run(f"cfg1_full_peaked_x{gain:g}", kw, shapes, batch=1, seed=200, q_gain=gain,
         lat_atol=LAT_ATOL if gain <= 8.0 else 3e-3)


def test_cfg1_full_batch4_vs_oracle():
    """The benchmarked batch (4 samples): batch-dependent split counts and grid shapes, float64 CUDA oracle."""
    kw, shapes = CFG["cfg1"]
    _run("cfg1_full_b4", kw, shapes, batch=4, seed=300)


@pytest.mark.parametrize("name,batch", [("cfg2", 8), ("cfg4", 4), ("cfg5", 2)])
def test_wsi_configs_full_size(name, batch):
    """cfg 2 (tab + 4096x1024, latent 256x512, batch 8), cfg 4 (omic + 8192x768, 4 per GPU), cfg 5 (65 536 x 512
    tokens): the generic K/V-projection streaming path at the full shapes."""
    kw, shapes = CFG[name]
    _run(f"{name}_full", kw, shapes, batch=batch, seed=400)


def test_wsi_full_size_peaked_attention():
    """cfg 5 token count with peaked logits on the generic path."""
    kw, shapes = CFG["cfg5"]
    _run("cfg5_full_peaked_x8", kw, shapes, batch=1, seed=500, q_gain=8.0)


def test_cfg3_full_width_depth_bf16():
    """cfg 3: latent 512x1024, depth 8, parameters and inputs in bf16 (one sample). Checked against the float64 oracle
    evaluated on the same bf16-rounded weights / inputs at the stated bf16 tolerance (the reference's own bf16 run
    differs from its fp32 run by 8.5e-3)."""
    kw, shapes = CFG["cfg3"]
    _run("cfg3_full_bf16", kw, shapes, batch=1, seed=600, bf16=True, rtol=1e-2, atol=2e-2, lat_atol=5e-2)
