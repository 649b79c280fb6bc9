"""Prints, per parameter, the error of the CUDA backward against the golden reference gradients (debug aid)."""
import json, os, sys
import numpy as np, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden, GOLDEN
from healnet_b200 import HealNet

for name in sys.argv[1:]:
    if name.startswith("grads_stream"):
        meta = json.load(open(os.path.join(GOLDEN, "index.json")))[name]
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
        xs = [torch.from_numpy(z[f"in/{i}"]).cuda() for i in range(meta["kwargs"]["n_modalities"])]
        mask = torch.from_numpy(z["in/mask"]).cuda() if "in/mask" in z.files else None
        g = z
    else:
        meta, sd, ins, outs, _ = load_golden(name)
        g = np.load(os.path.join(GOLDEN, f"grads_{name}.npz"))
        xs = [ins[str(i)].cuda() for i in range(meta["kwargs"]["n_modalities"])]
        mask = None
    model = HealNet(**meta["kwargs"]); model.load_state_dict(sd); model = model.cuda().train()
    logits = model(xs, mask=mask)
    loss = F.cross_entropy(logits, torch.from_numpy(g["targets"]).cuda())
    loss.backward()
    print(name, "loss", loss.item(), "want", float(g["loss"]))
    for k, p in model.named_parameters():
        want = torch.from_numpy(g["grad/" + k])
        got = p.grad.cpu()
        sc = float(want.abs().max())
        print("%-40s scale %.3e  err/scale %.3e" % (k, sc, float((got - want).abs().max()) / max(sc, 1e-20)))
