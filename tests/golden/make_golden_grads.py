"""Golden GRADIENTS for the backward pass (healnet/main.py:426-467): executes the
UNMODIFIED reference (loaded by file path, like make_golden.py) on the inputs and weights of existing forward fixtures
with autograd on, takes the cross-entropy of the logits against fixed targets (the reference's classification loss,
main.py:436-440) and stores d loss / d parameter for every parameter plus the loss value. The oracle's autograd is
pinned against these in tests/test_oracle_grad.py, so the CUDA backward of the next round has reference targets.
Run once in the build container: `python tests/golden/make_golden_grads.py`."""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference  # noqa: E402

CASES = ["tri_small", "omic_wsi_tied", "wide_heads"]


def main():
    ref = load_reference()
    torch.set_num_threads(4)
    index = json.load(open(os.path.join(HERE, "index.json")))
    for name in CASES:
        meta = index[name]
        z = np.load(os.path.join(HERE, name + ".npz"))
        sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
        xs = [torch.from_numpy(z[f"in/{i}"]) for i in range(meta["kwargs"]["n_modalities"])]
        model = ref.HealNet(**meta["kwargs"]).train()      # dropout 0: train == eval arithmetic
        model.load_state_dict(sd)
        batch, classes = xs[0].shape[0], meta["kwargs"]["out_dims"]
        targets = torch.arange(batch) % classes
        logits = model([t.clone() for t in xs])
        loss = F.cross_entropy(logits, targets)
        loss.backward()
        arrays = {"loss": np.asarray(loss.item(), dtype=np.float64), "targets": targets.numpy()}
        missing = []
        for k, p in model.named_parameters():
            if p.grad is None:
                missing.append(k)
                continue
            arrays["grad/" + k] = p.grad.numpy().copy()
        np.savez_compressed(os.path.join(HERE, f"grads_{name}.npz"), **arrays)
        print(name, "loss", float(loss.detach()), "grads", len(arrays) - 2, "params without grad", missing)


if __name__ == "__main__":
    sys.exit(main())
