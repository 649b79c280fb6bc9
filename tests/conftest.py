import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (NVIDIA B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this process")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """-> (meta, state_dict, inputs, outputs, extra) from tests/golden/<name>.npz (+ index.json)."""
    with open(os.path.join(GOLDEN, "index.json")) as f:
        meta = json.load(f)[name]
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    sd, ins, outs, extra = {}, {}, {}, {}
    for k in z.files:
        grp, key = k.split("/", 1)
        t = torch.from_numpy(z[k])
        if grp == "sd":
            sd[key] = t
        elif grp == "in":
            ins[key] = t
        elif grp == "out":
            outs[key] = t
        else:
            extra[k] = t
    return meta, sd, ins, outs, extra


@pytest.fixture(scope="session")
def golden():
    return load_golden
