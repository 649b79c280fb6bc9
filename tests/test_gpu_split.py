"""Token-axis sharding across GPUs on real devices: spawns one process per GPU (2 ranks)
running tools/run_split_check.py — sharded forward vs single-GPU forward AND vs the CPU oracle on the small-context,
generic and ragged + masked paths, bit-identical results on every rank, a rank delayed on the host (tolerated below
the exchange time-out, NaN + exception above it). Skipped on single-GPU boxes."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_token_sharded_forward_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "run_split_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("-> OK") == 4
