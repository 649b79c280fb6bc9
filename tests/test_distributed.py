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
