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


def pack_topk(idx: torch.Tensor, conf: torch.Tensor, out: torch.Tensor | None = None):
    """[n, k] int32 indices + [n, k] float32 confidences -> one [n, 2k] int32 buffer (confidences bit-cast), so the gather is
    ONE collective per step instead of two."""
    n, k = idx.shape
    if out is None:
        out = torch.empty((n, 2 * k), dtype=torch.int32, device=idx.device)
    out[:, :k] = idx
    out[:, k:] = conf.view(torch.int32)
    return out


def unpack_topk(packed: torch.Tensor):
    k = packed.shape[1] // 2
    return packed[:, :k], packed[:, k:].view(torch.float32)


def gather_topk_packed(packed: torch.Tensor, out: torch.Tensor, group=None):
    """Equal shards: all-gather [n, 2k] packed rows of every rank into out [world * n, 2k] (rank order = chunk order)."""
    dist.all_gather_into_tensor(out, packed, group=group)
    return out


def gather_topk(idx: torch.Tensor, conf: torch.Tensor, n_chunks: int, group=None):
    """All-gather ragged per-rank [n_r, k] results into chunk order [n_chunks, k] on every rank — ONE packed collective.

    Ranks own contiguous ranges (shard_range), so concatenating in rank order restores the global order.
    Shards are padded to the largest shard for the fixed-size collective and trimmed afterwards."""
    world = dist.get_world_size(group)
    k = idx.shape[1]
    per = (n_chunks + world - 1) // world
    pad = torch.zeros((per, 2 * k), dtype=torch.int32, device=idx.device)
    pack_topk(idx, conf, pad[: idx.shape[0]])
    allp = torch.empty((world * per, 2 * k), dtype=torch.int32, device=idx.device)
    gather_topk_packed(pad, allp, group=group)
    rows = []
    for r in range(world):
        lo, hi = shard_range(n_chunks, r, world)
        rows.append(allp[r * per: r * per + (hi - lo)])
    out = torch.cat(rows)
    oi, oc = unpack_topk(out)
    return oi.contiguous(), oc.contiguous()
