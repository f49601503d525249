"""CPU, world_size = 2, gloo: the N>1 host logic — contiguous sharding of the chunk index space and the top-k gather
(the only collective on the path).  The per-chunk "classifier" here is a deterministic stand-in so no GPU is needed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from birdnet_b200.dist import gather_topk, shard_range


def _fake_topk(chunk_ids, k=10):
    # deterministic per-chunk result: indices (7*id + j) % 6522, confidences descending
    ids = torch.as_tensor(chunk_ids, dtype=torch.int64)[:, None]
    j = torch.arange(k)[None, :]
    return ((7 * ids + j) % 6522).to(torch.int32), (1.0 / (1.0 + j + ids % 3)).to(torch.float32)


def _worker(rank, world, port, n_chunks, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_chunks, rank, world)
    idx, conf = _fake_topk(range(lo, hi))
    gi, gc = gather_topk(idx, conf, n_chunks)
    if rank == 0:
        out.put((gi.numpy(), gc.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("n_chunks", [79, 100, 1])
def test_sharded_topk_gather_matches_single_process(n_chunks):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_chunks, q)) for r in range(2)]
    for p in procs:
        p.start()
    gi, gc = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    wi, wc = _fake_topk(range(n_chunks))
    assert np.array_equal(gi, wi.numpy()) and np.array_equal(gc, wc.numpy())


def test_shard_ranges_partition_the_index_space():
    for n in (0, 1, 7, 79, 100000):
        for world in (1, 2, 3, 4, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
