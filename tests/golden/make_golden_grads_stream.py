"""Golden gradients for the STREAMING small-context path of the backward pass (token axes > 2048, context width <= 63:
image / volume modalities) — self-contained fixtures (kwargs in index.json; state_dict, inputs, optional mask, loss,
targets and d loss / d parameter in grads_stream_*.npz), produced by executing the UNMODIFIED reference
(/root/reference/healnet/models/healnet.py, loaded by file path) with autograd on, exactly like
make_golden_grads.py (cross-entropy of the logits, healnet/main.py:436-440).
Run once in the build container: `python tests/golden/make_golden_grads_stream.py`."""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference, randomise_affine  # noqa: E402

CASES = {
    # tab (generic precise path, N = 1) + image 48x48 (2304 tokens, C = 13) + volume 3x30x30 (2700 tokens, C = 18)
    "grads_stream_tri": dict(
        kwargs=dict(n_modalities=3, channel_dims=[30, 3, 3], num_spatial_axes=[1, 2, 3], out_dims=4, depth=2, l_c=40,
                    l_d=64, x_heads=2, l_heads=2, cross_dim_head=16, latent_dim_head=16),
        shapes=[(2, 1, 30), (2, 48, 48, 3), (2, 3, 30, 30, 3)], seed=31, masked=False, q_gain=1.0),
    # one image modality with a ragged last tile (50 x 60 = 3000 tokens), a token mask, a wide context (C = 35 + 10 = 45
    # -> 64-wide rows), peaked attention (to_q x 6) and the GELU gate
    "grads_stream_masked": dict(
        kwargs=dict(n_modalities=1, channel_dims=[35], num_spatial_axes=[2], out_dims=3, depth=2, l_c=70, l_d=48,
                    x_heads=3, l_heads=2, cross_dim_head=24, latent_dim_head=16, snn=False),
        shapes=[(2, 50, 60, 35)], seed=32, masked=True, q_gain=6.0),
}


def main():
    ref = load_reference()
    torch.set_num_threads(8)
    index_path = os.path.join(HERE, "index.json")
    index = json.load(open(index_path))
    for name, case in CASES.items():
        torch.manual_seed(case["seed"])
        gen = torch.Generator().manual_seed(case["seed"] + 1)
        model = ref.HealNet(**case["kwargs"]).train()
        randomise_affine(model, gen)
        if case["q_gain"] != 1.0:
            with torch.no_grad():
                for k, p in model.named_parameters():
                    if k.endswith("fn.to_q.weight") and ".norm_context" not in k and "layers" in k:
                        p.mul_(case["q_gain"])
        xs = [torch.rand(s, generator=gen) for s in case["shapes"]]
        mask = None
        if case["masked"]:
            n_tok = int(np.prod(case["shapes"][0][1:-1]))
            mask = torch.rand(case["shapes"][0][0], n_tok, generator=gen) > 0.35
            mask[:, :3] = True
        batch, classes = xs[0].shape[0], case["kwargs"]["out_dims"]
        targets = torch.arange(batch) % classes
        atts = [m for m in model.modules() if isinstance(m, ref.Attention)]
        logits = model([t.clone() for t in xs], mask=mask)
        assert all(a.attn_weights is not None for a in atts), "reference swallowed an exception"
        loss = F.cross_entropy(logits, targets)
        loss.backward()
        arrays = {"loss": np.asarray(loss.item(), dtype=np.float64), "targets": targets.numpy(),
                  "logits": logits.detach().numpy()}
        for k, v in model.state_dict().items():
            arrays["sd/" + k] = v.detach().numpy().copy()
        for i, t in enumerate(xs):
            arrays[f"in/{i}"] = t.numpy()
        if mask is not None:
            arrays["in/mask"] = mask.numpy()
        for k, p in model.named_parameters():
            assert p.grad is not None, k
            arrays["grad/" + k] = p.grad.numpy().copy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
        index[name] = dict(kwargs=case["kwargs"], shapes=[list(s) for s in case["shapes"]], masked=case["masked"])
        print(name, "loss", float(loss.detach()), "params", sum(1 for k in arrays if k.startswith("grad/")),
              "bytes", os.path.getsize(os.path.join(HERE, name + ".npz")))
    json.dump(index, open(index_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    sys.exit(main())
