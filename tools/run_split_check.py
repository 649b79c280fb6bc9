"""Token-axis sharding check, one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        tools/run_split_check.py [--bench]
Every rank runs (a) the plain single-GPU forward and (b) the token-sharded forward of the same inputs, checks that
(b) matches (a) to fp32 merge-order noise, that (b) matches the CPU ORACLE (oracle/healnet_oracle.py) at the north-star
tolerance, and that all ranks hold bit-identical results; then (c) one rank is held back on the host for a few seconds
before a sharded forward (the peers' combine kernels must wait, not merge stale partials) and (d) with a 1 s exchange
time-out the same delay must surface as NaN outputs + HealNetLibraryError, never as silently wrong logits.
Optionally times (a) against (b). Exit code 0 = all checks passed.
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from healnet_b200 import HealNet, HealNetLibraryError  # noqa: E402
from oracle import healnet_oracle as O  # noqa: E402  (the checker; test infrastructure only)

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
        cfg = O.OracleConfig(**{k: v for k, v in kw.items() if k in O.OracleConfig.__dataclass_fields__})
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        with torch.no_grad():
            want = O.forward(sd, cfg, [t.cpu() for t in xs], mask=mask.cpu() if mask is not None else None)
        oracle_ok = torch.allclose(got_log.cpu(), want, rtol=1e-3, atol=1e-4)
        err_oracle = (got_log.cpu() - want).abs().max().item()
        good = err_lat <= 2e-4 * max(scale, 1.0) and err_log <= 1e-4 and identical and oracle_ok
        ok = ok and good
        if rank == 0:
            print(f"{name}: sharded vs single-GPU max|err| latents {err_lat:.3e} (max |ref| {scale:.2f}) logits "
                  f"{err_log:.3e}; sharded vs CPU oracle logits {err_oracle:.3e}; ranks bit-identical: {identical} "
                  f"-> {'OK' if good else 'FAIL'}", flush=True)
        if name == "tri_small_ctx":
            import time
            # (c) host-side skew: the last rank enters the forward 3 s late; everybody must still get the right result
            with torch.no_grad():
                model.enable_token_sharding(min_tokens=2049)
                model(list(xs))
                torch.cuda.synchronize()
                dist.barrier()
                if rank == world - 1:
                    time.sleep(3.0)
                late = model(list(xs))
                torch.cuda.synchronize()
                model.check_token_sharding()
                skew_ok = torch.equal(late, got_log)
                # (d) same skew with a 1 s time-out: the early ranks must get NaN + an exception, not stale numbers
                model.disable_token_sharding()
                model.token_sharding_timeout_s = 1.0
                dist.barrier()               # nobody unmaps a buffer a peer's kernels may still be reading
                model._release_exchange()
                model.enable_token_sharding(min_tokens=2049)
                model(list(xs))
                torch.cuda.synchronize()
                dist.barrier()
                if rank == world - 1:
                    time.sleep(4.0)
                out = model(list(xs))
                torch.cuda.synchronize()
                raised = False
                try:
                    model.check_token_sharding()
                except HealNetLibraryError:
                    raised = True
                timeout_ok = (rank == world - 1) or (raised and bool(torch.isnan(out).all()))
                model.disable_token_sharding()
                model.token_sharding_timeout_s = 30.0
                dist.barrier()
                model._release_exchange()
            flags = torch.tensor([int(skew_ok), int(timeout_ok)], device="cuda")
            dist.all_reduce(flags, op=dist.ReduceOp.MIN)
            good2 = bool(flags[0].item()) and bool(flags[1].item())
            ok = ok and good2
            if rank == 0:
                print(f"delayed rank: 3 s host skew tolerated = {bool(flags[0].item())}; 1 s time-out surfaces as NaN + "
                      f"HealNetLibraryError = {bool(flags[1].item())} -> {'OK' if good2 else 'FAIL'}", flush=True)
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
