"""Token-axis sharding check, one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        tools/run_split_check.py [--bench]
Every rank runs (a) the plain single-GPU forward and (b) the token-sharded forward of the same inputs, checks that
(b) matches (a) to fp32 merge-order noise and that all ranks hold bit-identical results, and optionally times both.
Exit code 0 = all checks passed.
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from healnet_b200 import HealNet  # noqa: E402

CASES = {
    # name: (constructor kwargs, input shapes (without batch), batch)
    "tri_small_ctx": (dict(n_modalities=3, channel_dims=[200, 3, 3], num_spatial_axes=[1, 2, 3], out_dims=4, l_c=128,
                           l_d=128, depth=2), [(1, 200), (96, 96, 3), (6, 64, 64, 3)], 2),
    "wsi_generic": (dict(n_modalities=2, channel_dims=[300, 256], num_spatial_axes=[1, 1], out_dims=3, l_c=256,
                         l_d=256, depth=2), [(1, 300), (6000, 256)], 2),
    "ragged_masked": (dict(n_modalities=1, channel_dims=[3], num_spatial_axes=[2], out_dims=2, l_c=128, l_d=128,
                           depth=1), [(70, 71, 3)], 3),
}
BENCH = {
    "cfg1": (dict(n_modalities=3, channel_dims=[2000, 3, 3], num_spatial_axes=[1, 2, 3], out_dims=4, l_c=512, l_d=512),
             [(1, 2000), (224, 224, 3), (12, 224, 224, 3)]),
}


def timed(fn, steps=8, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bench", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    ok = True
    for name, (kw, shapes, batch) in CASES.items():
        torch.manual_seed(0)
        model = HealNet(**kw).eval().cuda()
        g = torch.Generator().manual_seed(1)
        xs = [torch.rand(batch, *s, generator=g).cuda() for s in shapes]
        mask = None
        if name == "ragged_masked":
            mask = (torch.rand(batch, 70 * 71, generator=g) > 0.3).cuda()
        with torch.no_grad():
            ref_lat = model(list(xs), mask=mask, return_embeddings=True)
            ref_log = model(list(xs), mask=mask)
            model.enable_token_sharding(min_tokens=2049)
            got_lat = model(list(xs), mask=mask, return_embeddings=True)
            got_log = model(list(xs), mask=mask)
            model.disable_token_sharding()
        torch.cuda.synchronize()
        err_lat = (got_lat - ref_lat).abs().max().item()
        err_log = (got_log - ref_log).abs().max().item()
        gathered = [torch.empty_like(got_lat) for _ in range(world)]
        dist.all_gather(gathered, got_lat.contiguous())
        identical = all(torch.equal(gathered[0], t) for t in gathered)
        scale = ref_lat.abs().max().item()
        good = err_lat <= 2e-4 * max(scale, 1.0) and err_log <= 1e-4 and identical
        ok = ok and good
        if rank == 0:
            print(f"{name}: sharded vs single-GPU max|err| latents {err_lat:.3e} (max |ref| {scale:.2f}) logits "
                  f"{err_log:.3e}; ranks bit-identical: {identical} -> {'OK' if good else 'FAIL'}", flush=True)
        del model
    if args.bench:
        for name, (kw, shapes) in BENCH.items():
            for batch in (1, 4):
                torch.manual_seed(0)
                model = HealNet(**kw).eval().cuda()
                xs = [torch.rand(batch, *s).cuda() for s in shapes]
                with torch.no_grad():
                    t_one = timed(lambda: model(list(xs)))
                    model.enable_token_sharding(min_tokens=8192)
                    t_split = timed(lambda: model(list(xs)))
                if rank == 0:
                    print(f"{name} batch {batch}: single GPU {t_one:.3f} ms/forward ({batch / t_one * 1e3:.1f} samples/s)"
                          f" | token-sharded over {world} GPUs {t_split:.3f} ms ({batch / t_split * 1e3:.1f} samples/s,"
                          f" x{t_one / t_split:.2f})", flush=True)
                del model
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
