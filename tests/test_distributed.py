"""CPU, gloo, world_size 2: the batch-sharding + single all-gather host logic of healnet_b200.distributed,
with the oracle standing in for the per-rank GPU compute."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from healnet_b200.distributed import gather_rows, shard_bounds, sharded_forward


def test_shard_bounds_partition_the_batch():
    for batch in (1, 2, 5, 8, 33):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import healnet_oracle as O
    from conftest import load_golden
    meta, sd, ins, outs, _ = load_golden("tri_small")
    kw = {k: v for k, v in meta["kwargs"].items() if k in O.OracleConfig.__dataclass_fields__}
    cfg = O.OracleConfig(**kw)
    g = torch.Generator().manual_seed(123)
    tensors = [torch.rand((batch,) + tuple(s[1:]), generator=g) for s in meta["shapes"]]
    compute = lambda ts, **k: O.forward(sd, cfg, list(ts), **k)
    full = compute(tensors)
    got = sharded_forward(compute, tensors)
    torch.testing.assert_close(got, full, rtol=1e-5, atol=1e-6)
    emb = sharded_forward(compute, tensors, return_embeddings=True)
    assert emb.shape[0] == batch
    torch.testing.assert_close(emb, compute(tensors, return_embeddings=True), rtol=1e-5, atol=1e-6)
    # missing modality stays missing on every rank
    miss = [tensors[0], None, tensors[2]]
    torch.testing.assert_close(sharded_forward(compute, miss), compute(miss), rtol=1e-5, atol=1e-6)
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(result_dir, f"ok{rank}"), "w").close()


@pytest.mark.parametrize("batch", [4, 3, 1])
def test_sharded_forward_world2(tmp_path, batch):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, batch, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


# ---------------------------------------------------------------------------------- token-axis sharding (row f4)
from healnet_b200.distributed import merge_softmax_partials, token_shard_bounds  # noqa: E402


def test_token_shard_bounds_cover_the_axis_in_tile_units():
    for n in (2049, 4970, 50176, 602112, 65536):
        for world in (1, 2, 3, 4, 8):
            spans = [token_shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and b > a and b % 64 == 0      # cuts on tile boundaries, nobody empty
            tiles = [-(-(b - a) // 64) for a, b in spans]
            assert max(tiles) - min(tiles) <= 1


def _split_worker(rank, world, port, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(7)
    b, h, L, N, dh = 2, 3, 16, 4970, 8
    q = torch.randn(b, h, L, dh, generator=g, dtype=torch.float64)
    k = torch.randn(b, h, N, dh, generator=g, dtype=torch.float64)
    v = torch.randn(b, h, N, dh, generator=g, dtype=torch.float64)
    keep = torch.rand(b, N, generator=g) > 0.2
    sim = torch.einsum("bhld,bhnd->bhln", q, k).masked_fill(~keep[:, None, None, :], float("-inf"))
    full = torch.softmax(sim, dim=-1) @ v
    lo, hi = token_shard_bounds(N, world, rank)
    s = sim[..., lo:hi]
    m = s.amax(dim=-1)                                   # what a rank's streaming kernel leaves behind:
    p = torch.exp(s - m[..., None])                      # running max, row sum and un-normalised accumulator
    part = (m, p.sum(-1), p @ v[:, :, lo:hi])
    gathered = [[torch.empty_like(t) for _ in range(world)] for t in part]
    for t, outs in zip(part, gathered):
        dist.all_gather(outs, t.contiguous())
    merged = merge_softmax_partials(*gathered)
    torch.testing.assert_close(merged, full, rtol=1e-10, atol=1e-12)
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(result_dir, f"ok{rank}"), "w").close()


def test_token_sharded_partials_merge_world2(tmp_path):
    world = 2
    mp.spawn(_split_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
