"""CPU, world_size = 2, gloo: the N>1 host logic — contiguous sharding of the chunk index space and the top-k gather
(the only collective on the path).  The per-chunk "classifier" here is a deterministic stand-in so no GPU is needed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from birdnet_b200.dist import gather_topk, gather_topk_packed, pack_topk, shard_range, unpack_topk


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


def _worker_packed(rank, world, port, n_per_rank, out):
    """bench.py's N > 1 step: equal shards, ONE packed [n, 2k] int32 all-gather (confidences bit-cast), double-buffered."""
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    got = []
    packed = [torch.empty((n_per_rank, 20), dtype=torch.int32) for _ in range(2)]
    gathered = [torch.empty((world * n_per_rank, 20), dtype=torch.int32) for _ in range(2)]
    for step in range(3):
        ids = range(step * 1000 + rank * n_per_rank, step * 1000 + (rank + 1) * n_per_rank)
        idx, conf = _fake_topk(ids)
        j = step & 1
        gather_topk_packed(pack_topk(idx, conf, packed[j]), gathered[j])
        gi, gc = unpack_topk(gathered[j])
        got.append((gi.clone().numpy(), gc.clone().numpy()))
    if rank == 0:
        out.put(got)
    dist.barrier()
    dist.destroy_process_group()


def test_packed_gather_is_what_two_gathers_would_give():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n = 37
    procs = [ctx.Process(target=_worker_packed, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for step, (gi, gc) in enumerate(got):
        wi, wc = _fake_topk(range(step * 1000, step * 1000 + 2 * n))
        assert np.array_equal(gi, wi.numpy()) and np.array_equal(gc, wc.numpy())       # rank order = chunk order, bits intact


def test_pack_unpack_round_trip_keeps_the_bits():
    idx, conf = _fake_topk(range(50))
    conf[3, 2] = float("nan"); conf[4, 0] = -0.0
    p = pack_topk(idx, conf)
    assert p.dtype == torch.int32 and p.shape == (50, 20)
    i2, c2 = unpack_topk(p)
    assert torch.equal(i2, idx) and torch.equal(c2.view(torch.int32), conf.view(torch.int32))


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
