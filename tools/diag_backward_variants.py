"""Debug aid: CUDA backward vs the oracle's autograd (CPU) on model variants; prints the worst per-tensor error."""
import os, sys
import torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from healnet_b200 import HealNet
from oracle import healnet_oracle as O

def run(tag, kw, shapes, seed=0, verbose=False):
    torch.manual_seed(seed)
    model = HealNet(**kw)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("norm.weight") or n.endswith("norm_context.weight") or n == "to_logits.1.weight":
                p.copy_(1.0 + 0.5 * torch.randn(p.shape, generator=g))
            elif n.endswith(".bias"):
                p.copy_(0.5 * torch.randn(p.shape, generator=g))
    xs = [torch.rand(s, generator=g) for s in shapes]
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    cfg = O.OracleConfig(**{k: v for k, v in kw.items() if k in O.OracleConfig.__dataclass_fields__})
    y = torch.arange(shapes[0][0]) % kw["out_dims"]
    F.cross_entropy(O.forward(sd, cfg, xs), y).backward()
    model = model.cuda().train()
    F.cross_entropy(model([t.cuda() for t in xs]), y.cuda()).backward()
    worst = (0.0, None)
    # tied aliases: sum grads over identical tensors
    groups = {}
    for k, v in sd.items():
        groups.setdefault((tuple(v.shape), v.detach().numpy().tobytes()), []).append(k)
    alias = {k: ks for ks in groups.values() for k in ks}
    for k, p in model.named_parameters():
        want = sum(sd[a].grad for a in alias[k] if sd[a].grad is not None)
        sc = float(want.abs().max()) if torch.is_tensor(want) else 0.0
        if sc < 1e-7:
            continue
        e = float((p.grad.cpu() - want).abs().max()) / sc
        if verbose:
            print("   %-36s %.2e" % (k, e))
        if e > worst[0]:
            worst = (e, k)
    print("%-28s worst %.2e  %s" % (tag, worst[0], worst[1]))

base = dict(channel_dims=[70, 96], cross_dim_head=27, depth=3, l_c=17, l_d=40, l_heads=2, latent_dim_head=8, max_freq=6.0,
            n_modalities=2, num_freq_bands=3, num_spatial_axes=[1, 1], out_dims=4, snn=False, x_heads=1)
sh = [(2, 1, 70), (2, 150, 96)]
run("depth3 untied", base, sh)
run("depth2 untied", dict(base, depth=2), sh)
run("depth3 tied", dict(base, weight_tie_layers=True), sh)
run("depth2 tied", dict(base, depth=2, weight_tie_layers=True), sh)
run("depth3 tied snn", dict(base, weight_tie_layers=True, snn=True), sh)
run("depth3 tied seed5", dict(base, weight_tie_layers=True), sh, seed=5)
run("depth3 tied tab only", dict(base, weight_tie_layers=True, n_modalities=1, channel_dims=[70], num_spatial_axes=[1]), sh[:1])
run("depth3 tied wsi only", dict(base, weight_tie_layers=True, n_modalities=1, channel_dims=[96], num_spatial_axes=[1]), sh[1:])
run("depth3 tied verbose", dict(base, weight_tie_layers=True), sh, verbose=True)
