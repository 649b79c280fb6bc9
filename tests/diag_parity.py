"""Prints error statistics of the CUDA path vs the golden fixtures / oracle (run on the GPU box)."""
import sys, os, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import load_golden
from healnet_b200 import HealNet, Attention
from oracle import healnet_oracle as O


def stats(name, got, want):
    got, want = got.double().cpu(), want.double().cpu()
    err = (got - want).abs()
    tol = 1e-4 + 1e-3 * want.abs()
    print(f"{name:44s} max|err|={err.max():.3e} rms_err={err.pow(2).mean().sqrt():.3e} max|ref|={want.abs().max():.3e} "
          f"rms_ref={want.pow(2).mean().sqrt():.3e} viol(1e-3/1e-4)={(err > tol).float().mean()*100:.2f}% "
          f"worst err/tol={(err / tol).max():.2f}", flush=True)


for name in ["tri_small", "omic_wsi_tied", "plain_no_head", "two_ltiles", "masked"]:
    meta, sd, ins, outs, _ = load_golden(name)
    m = HealNet(**meta["kwargs"]); m.load_state_dict(sd); m = m.cuda().eval()
    x = [ins[str(i)].cuda() for i in range(meta["kwargs"]["n_modalities"])]
    kw = {"mask": ins["mask"].cuda()} if "mask" in ins else {}
    stats(name + " latents", m(list(x), return_embeddings=True, **kw), outs["latents"])
    if "logits" in outs:
        stats(name + " logits", m(list(x), **kw), outs["logits"])
    if "missing1_latents" in outs:
        miss = [x[0], None] + x[2:]
        stats(name + " missing latents", m(miss, return_embeddings=True), outs["missing1_latents"])
        stats(name + " missing verbose latents", m(miss, return_embeddings=True, verbose=True), outs["missing1_verbose_latents"])
        stats(name + " short list latents", m([x[0]], return_embeddings=True), outs["short_list_latents"])

meta, sd, ins, outs, extra = load_golden("attention")
att = Attention(**meta["kwargs"]); att.load_state_dict(sd); att = att.cuda().eval()
stats("attention cross", att(ins["x"].cuda(), context=ins["context"].cuda()), outs["cross"])
stats("attention cross masked", att(ins["x"].cuda(), context=ins["context"].cuda(), mask=ins["mask"].cuda()), outs["cross_masked"])

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_parity import ORACLE_CASES, _randomise, _cfg
for name, (kw, shapes) in ORACLE_CASES.items():
    torch.manual_seed(11)
    model = HealNet(**kw).eval(); _randomise(model, 5)
    g = torch.Generator().manual_seed(3)
    xs = [torch.rand(s, generator=g) for s in shapes]
    sdd = {k: v.clone() for k, v in model.state_dict().items()}
    want = O.forward(sdd, _cfg(kw), xs); want_lat = O.forward(sdd, _cfg(kw), xs, return_embeddings=True)
    want64 = O.forward(sdd, _cfg(kw), xs, dtype=torch.float64)
    model.cuda()
    stats(name + " logits", model([t.cuda() for t in xs]), want)
    stats(name + " latents", model([t.cuda() for t in xs], return_embeddings=True), want_lat)
    stats(name + " oracle fp32 vs fp64 logits", want, want64)
