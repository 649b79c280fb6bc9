"""Generates the golden fixtures in this directory by executing the UNMODIFIED reference
(/root/reference/healnet/models/healnet.py, loaded by file path) on seeded inputs and weights.

Run once in the build container (`python tests/golden/make_golden.py`); the resulting `*.npz` files are
committed because /root/reference does not exist on the GPU box. Each fixture holds the constructor
kwargs (JSON), the full reference state_dict, the inputs, and the reference outputs (logits and
`return_embeddings=True` latents). LayerNorm affines and all biases are re-randomised after construction:
the default init (gamma=1, beta=0) would hide affine-folding bugs.
"""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference/healnet/models/healnet.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_healnet", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def randomise_affine(model, gen):
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("norm.weight") or name.endswith("norm_context.weight") or name == "to_logits.1.weight":
                p.copy_(1.0 + 0.5 * torch.randn(p.shape, generator=gen))
            elif name.endswith(".bias"):
                p.copy_(0.5 * torch.randn(p.shape, generator=gen))


def run_ref(ref, model, tensors, **kw):
    """Fresh list per call (the reference mutates it, healnet.py:222); counts the attention calls that really
    ran (the reference swallows exceptions, :238-239)."""
    atts = [m for m in model.modules() if isinstance(m, ref.Attention)]
    for a in atts:
        a.attn_weights = None
    with torch.no_grad():
        out = model([None if t is None else t.clone() for t in tensors], **kw)
    ran = sum(a.attn_weights is not None for a in atts)
    for a in atts:
        a.attn_weights = None
    return out, ran


CASES = {
    # 3 modalities like the README example, tiny: tab (C=45 -> 64-wide small path), img (C=13), vol (C=18)
    "tri_small": dict(
        kwargs=dict(n_modalities=3, channel_dims=[40, 3, 3], num_spatial_axes=[1, 2, 3], out_dims=4, depth=2,
                    l_c=24, l_d=32, x_heads=2, l_heads=2, cross_dim_head=16, latent_dim_head=16),
        shapes=[(3, 1, 40), (3, 12, 10, 3), (3, 3, 6, 5, 3)], seed=1),
    # production-like: one odd-sized cross head, wide contexts (generic K/V-projection path), tied layers, GELU gate
    "omic_wsi_tied": dict(
        kwargs=dict(n_modalities=2, channel_dims=[70, 96], num_spatial_axes=[1, 1], out_dims=4, depth=3,
                    l_c=17, l_d=40, x_heads=1, l_heads=2, cross_dim_head=27, latent_dim_head=8,
                    weight_tie_layers=True, snn=False, num_freq_bands=3, max_freq=6.0),
        shapes=[(2, 1, 70), (2, 150, 96)], seed=2),
    # no Fourier features, no head, no latent self-attention
    "plain_no_head": dict(
        kwargs=dict(n_modalities=2, channel_dims=[10, 50], num_spatial_axes=[1, 2], out_dims=3, depth=2,
                    l_c=16, l_d=24, x_heads=2, l_heads=2, cross_dim_head=8, latent_dim_head=8,
                    fourier_encode_data=False, final_classifier_head=False, self_per_cross_attn=0),
        shapes=[(2, 5, 10), (2, 9, 7, 50)], seed=3),
    # more than one 128-row latent tile, more than one 64-token tile with a ragged tail, default head sizes
    # tuned production hyper-parameters of the reference (config/best_hyperparams.yml, ucec): one cross head of
    # 103 dims (> 64: two 64-column atoms per head), tiny odd latent array, no latent self-attention (so the
    # latent_dim_head of 51 is unused), max_freq 2
    "prod_ucec": dict(
        kwargs=dict(n_modalities=2, channel_dims=[120, 80], num_spatial_axes=[1, 1], out_dims=4, depth=2,
                    l_c=16, l_d=65, x_heads=1, l_heads=8, cross_dim_head=103, latent_dim_head=51,
                    self_per_cross_attn=0, max_freq=2.0),
        shapes=[(3, 1, 120), (3, 90, 80)], seed=5),
    # wide heads on both sides: cross 2 x 80, latent 2 x 100 (kirp / blca use latent_dim_head 113 / 127)
    "wide_heads": dict(
        kwargs=dict(n_modalities=2, channel_dims=[30, 3], num_spatial_axes=[1, 2], out_dims=3, depth=2,
                    l_c=40, l_d=48, x_heads=2, l_heads=2, cross_dim_head=80, latent_dim_head=100),
        shapes=[(2, 4, 30), (2, 9, 11, 3)], seed=6),
    "two_ltiles": dict(
        kwargs=dict(n_modalities=2, channel_dims=[5, 3], num_spatial_axes=[1, 2], out_dims=2, depth=1,
                    l_c=130, l_d=64, x_heads=2, l_heads=2, cross_dim_head=64, latent_dim_head=64),
        shapes=[(2, 3, 5), (2, 31, 29, 3)], seed=4),
}


def main():
    ref = load_reference()
    torch.set_num_threads(4)
    index = {}
    for name, case in CASES.items():
        gen = torch.Generator().manual_seed(case["seed"])
        torch.manual_seed(case["seed"])
        model = ref.HealNet(**case["kwargs"]).eval()
        init_sig = {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in model.state_dict().items()}
        randomise_affine(model, gen)
        tensors = [torch.rand(s, generator=gen) for s in case["shapes"]]
        arrays = {}
        for k, v in model.state_dict().items():
            arrays["sd/" + k] = v.numpy().copy()
        for i, t in enumerate(tensors):
            arrays[f"in/{i}"] = t.numpy().copy()
        n_att = sum(isinstance(m, ref.Attention) for m in model.modules())
        has_head = case["kwargs"].get("final_classifier_head", True)
        out, ran = run_ref(ref, model, tensors)
        arrays["out/logits" if has_head else "out/latents"] = out.numpy()
        emb, _ = run_ref(ref, model, tensors, return_embeddings=True)
        arrays["out/latents"] = emb.numpy()
        meta = dict(kwargs=case["kwargs"], shapes=case["shapes"], seed=case["seed"], init_sig=init_sig,
                    attention_modules=n_att, attention_calls_ran=ran)
        # missing-modality semantics (healnet.py:229-245): default skips only the cross block; verbose skips both
        if case["kwargs"]["n_modalities"] >= 2:
            miss = [tensors[0]] + [None] + list(tensors[2:])
            o1, ran1 = run_ref(ref, model, miss, return_embeddings=True)
            arrays["out/missing1_latents"] = o1.numpy()
            o2, ran2 = run_ref(ref, model, miss, return_embeddings=True, verbose=True)
            arrays["out/missing1_verbose_latents"] = o2.numpy()
            short = [tensors[0]]  # shorter list: IndexError swallowed for the rest (main.py:526-541)
            o3, ran3 = run_ref(ref, model, short, return_embeddings=True)
            arrays["out/short_list_latents"] = o3.numpy()
            meta.update(missing_calls=[ran1, ran2, ran3])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
        index[name] = meta
        print(name, "attention calls ran", ran, "of", n_att, "params", sum(p.numel() for p in model.parameters()))

    # masked single-modality case: mask (b, N) bool applied to the cross-attention (healnet.py:411-415)
    gen = torch.Generator().manual_seed(7)
    torch.manual_seed(7)
    kw = dict(n_modalities=1, channel_dims=[20], num_spatial_axes=[1], out_dims=3, depth=2, l_c=12, l_d=32,
              x_heads=2, l_heads=2, cross_dim_head=16, latent_dim_head=16)
    model = ref.HealNet(**kw).eval()
    randomise_affine(model, gen)
    x = torch.rand((3, 100, 20), generator=gen)
    mask = torch.rand((3, 100), generator=gen) > 0.4
    mask[2, 64:] = False  # a whole 64-token tile masked out
    arrays = {"sd/" + k: v.numpy().copy() for k, v in model.state_dict().items()}
    arrays["in/0"] = x.numpy()
    arrays["in/mask"] = mask.numpy()
    out, ran = run_ref(ref, model, [x], mask=mask)
    arrays["out/logits"] = out.numpy()
    emb, _ = run_ref(ref, model, [x], mask=mask, return_embeddings=True)
    arrays["out/latents"] = emb.numpy()
    np.savez_compressed(os.path.join(HERE, "masked.npz"), **arrays)
    index["masked"] = dict(kwargs=kw, shapes=[[3, 100, 20]], seed=7, attention_calls_ran=ran)
    print("masked: attention calls ran", ran)

    # stand-alone Attention (healnet.py:369-426), shapes after the reference's test_attention, smaller
    gen = torch.Generator().manual_seed(9)
    torch.manual_seed(9)
    att = ref.Attention(query_dim=32, context_dim=77, heads=4, dim_head=24).eval()
    lat = torch.randn((3, 40, 32), generator=gen)
    ctx = torch.randn((3, 70, 77), generator=gen)
    am = torch.rand((3, 70), generator=gen) > 0.3
    arrays = {"sd/" + k: v.numpy().copy() for k, v in att.state_dict().items()}
    arrays.update({"in/x": lat.numpy(), "in/context": ctx.numpy(), "in/mask": am.numpy()})
    with torch.no_grad():
        arrays["out/cross"] = att(lat, context=ctx).numpy()
        arrays["out/cross_masked"] = att(lat, context=ctx, mask=am).numpy()
        arrays["attn/cross"] = att.attn_weights.numpy().copy()
    att2 = ref.Attention(query_dim=32, heads=2, dim_head=16).eval()
    arrays.update({"sd2/" + k: v.numpy().copy() for k, v in att2.state_dict().items()})
    with torch.no_grad():
        arrays["out/self"] = att2(lat).numpy()
    np.savez_compressed(os.path.join(HERE, "attention.npz"), **arrays)
    index["attention"] = dict(kwargs=dict(query_dim=32, context_dim=77, heads=4, dim_head=24),
                              kwargs_self=dict(query_dim=32, heads=2, dim_head=16))

    # Fourier feature table (healnet.py:211-217, 292-302) for a few axis sizes incl. the size-1 edge case
    four = {}
    for size in (1, 2, 7, 224):
        pos = torch.linspace(-1., 1., steps=size)
        four[f"fourier/{size}"] = ref.fourier_encode(pos, 10.0, 2).numpy()
    np.savez_compressed(os.path.join(HERE, "fourier.npz"), **four)

    with open(os.path.join(HERE, "index.json"), "w") as f:
        json.dump(index, f, indent=1, sort_keys=True)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    sys.exit(main())
