#!/usr/bin/env python
"""Is hn_forward capturable in a CUDA graph, and what does replaying it buy at small batch? (the forward is a pure
stream-ordered launch sequence: no host sync, no allocation). Prints ms per forward, eager vs graph replay, and
checks that the replayed logits equal the eager ones.   python tools/graph_forward.py [batch ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from healnet_b200 import HealNet  # noqa: E402

kw, shapes, _ = bench.WORKLOADS["cfg1"]
torch.manual_seed(0)
model = HealNet(**kw).eval().cuda()
for b in [int(a) for a in sys.argv[1:]] or [1, 4]:
    xs = [torch.rand((b,) + tuple(s), device="cuda") for s in shapes]
    with torch.no_grad():
        for _ in range(3):
            want = model(list(xs))
        torch.cuda.synchronize()

        def timed(fn, n=20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n

        eager = timed(lambda: model(list(xs)))
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            model(list(xs))                      # warm-up on the capture stream
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            got = model(list(xs))
        g.replay()
        torch.cuda.synchronize()
        ok = torch.equal(got, want)
        replay = timed(g.replay)
    print(f"batch {b}: eager {eager:.3f} ms, graph replay {replay:.3f} ms per forward, identical logits: {ok}")
