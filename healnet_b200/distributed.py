"""Batch-sharded multi-GPU forward: one process per GPU, contiguous batch slices per rank, parameters
replicated, and ONE all-gather of the (b_local, out_dims) logits at the end. The
reference has no distributed code at all; every op of the path is per-sample, so
no other collective is needed.

`torch.distributed` is plumbing only: NCCL over NVLink on the GPU box, gloo in the CPU tests that cover
this host logic with a stand-in compute function.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of the batch owned by `rank`; the first `batch % world_size` ranks get one
    extra sample."""
    base, extra = divmod(batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def token_shard_bounds(n_tokens: int, world_size: int, rank: int, align: int = 64) -> Tuple[int, int]:
    """Token-axis sharding: contiguous slice [lo, hi) of a modality's flattened token axis
    owned by `rank`, cut at multiples of `align` tokens (the streaming kernel's tile) so only the last rank sees a
    ragged tile. Slices may be empty only when there are fewer tiles than ranks."""
    tiles = -(-n_tokens // align)
    base, extra = divmod(tiles, world_size)
    t_lo = rank * base + min(rank, extra)
    t_hi = t_lo + base + (1 if rank < extra else 0)
    return min(t_lo * align, n_tokens), min(t_hi * align, n_tokens)


def merge_softmax_partials(maxes, sums, accs) -> torch.Tensor:
    """Host restatement of what the combine kernels do with the ranks' partials of one attention row block
    (rowops.cu: merge_signal_kernel / combine_*_kernel): partial r holds the running max m_r, the row sum l_r and the
    un-normalised accumulator A_r of softmax(s) v over rank r's tokens; merged in rank order so every rank gets the
    same bits.  out = sum_r e^(m_r - M) A_r / sum_r e^(m_r - M) l_r,  M = max_r m_r."""
    big = torch.stack(list(maxes)).amax(dim=0)
    num, den = 0, 0
    for m, l, a in zip(maxes, sums, accs):
        w = torch.where(torch.isinf(m) & (m < 0), torch.zeros_like(m), torch.exp(m - big))
        num = num + w[..., None] * a
        den = den + w * l
    return num / den[..., None]


def shard_batch(tensors: Sequence[Optional[torch.Tensor]], world_size: int, rank: int):
    """Slices every present modality (and nothing else) along the batch axis."""
    batch = next(t.shape[0] for t in tensors if t is not None)
    lo, hi = shard_bounds(batch, world_size, rank)
    return [None if t is None else t[lo:hi] for t in tensors], (lo, hi, batch)


def gather_rows(local: torch.Tensor, batch: int, group=None) -> torch.Tensor:
    """All-gathers per-rank row blocks (uneven shards allowed) into the full (batch, ...) tensor on every rank."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    per = -(-batch // world)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    pieces = []
    for r in range(world):
        lo, hi = shard_bounds(batch, world, r)
        pieces.append(out[r * per: r * per + (hi - lo)])
    return torch.cat(pieces, dim=0)


def sharded_forward(compute: Callable[[List[Optional[torch.Tensor]]], torch.Tensor],
                    tensors: Sequence[Optional[torch.Tensor]], mask: Optional[torch.Tensor] = None,
                    group=None, **kwargs) -> torch.Tensor:
    """Runs `compute` (normally a `healnet_b200.HealNet` on this rank's GPU) on this rank's batch slice and
    returns the full-batch result on every rank. `tensors` hold the GLOBAL batch (host or device)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    local, (lo, hi, batch) = shard_batch(tensors, world, rank)
    if mask is not None:
        kwargs["mask"] = mask[lo:hi]
    if hi > lo:
        out = compute(local, **kwargs)
    else:  # more ranks than samples: contribute an empty block of the right trailing shape
        probe = compute([None if t is None else t[:1] for t in tensors], **kwargs)
        out = probe[:0]
    return gather_rows(out, batch, group) if world > 1 else out
