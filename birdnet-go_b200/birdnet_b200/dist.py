"""Multi-GPU plumbing for the embarrassingly parallel chunk batch.

Chunks are independent, every rank holds a full weight replica, and the ONLY collective is the gather of the
per-chunk top-k (k int32 indices + k float32 confidences = 80 B/chunk at k = 10).  One process per GPU
(torchrun); NCCL on GPUs, gloo in the CPU tests."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_chunks: int, rank: int, world: int):
    """Contiguous range [lo, hi) of the chunk index space owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_chunks, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_topk(idx: torch.Tensor, conf: torch.Tensor, n_chunks: int, group=None):
    """All-gather ragged per-rank [n_r, k] results into chunk order [n_chunks, k] on every rank.

    Ranks own contiguous ranges (shard_range), so concatenating in rank order restores the global order.
    Shards are padded to the largest shard for the fixed-size collective and trimmed afterwards."""
    world = dist.get_world_size(group)
    k = idx.shape[1]
    per = (n_chunks + world - 1) // world
    pad_i = torch.zeros((per, k), dtype=idx.dtype, device=idx.device)
    pad_c = torch.zeros((per, k), dtype=conf.dtype, device=conf.device)
    pad_i[: idx.shape[0]] = idx
    pad_c[: conf.shape[0]] = conf
    all_i = torch.empty((world * per, k), dtype=idx.dtype, device=idx.device)
    all_c = torch.empty((world * per, k), dtype=conf.dtype, device=conf.device)
    dist.all_gather_into_tensor(all_i, pad_i, group=group)
    dist.all_gather_into_tensor(all_c, pad_c, group=group)
    out_i, out_c = [], []
    for r in range(world):
        lo, hi = shard_range(n_chunks, r, world)
        out_i.append(all_i[r * per: r * per + (hi - lo)])
        out_c.append(all_c[r * per: r * per + (hi - lo)])
    return torch.cat(out_i), torch.cat(out_c)
