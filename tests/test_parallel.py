"""N>1 host logic on CPU: two gloo ranks shard the tracker streams, receive the weight blob by ONE broadcast,
agree on max-over-ranks timing, and gather per-stream records in stream order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from siammask_b200.parallel import (broadcast_weights, gather_stream_records, max_over_ranks, owner_of,
                                    shard_streams)


def test_shards_partition_the_streams():
    for n, w in [(512, 8), (64, 1), (10, 4), (3, 8), (0, 2)]:
        seen = []
        for r in range(w):
            rg = shard_streams(n, w, r)
            seen += list(rg)
        assert seen == list(range(n))
        sizes = [len(shard_streams(n, w, r)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1
    assert len(shard_streams(512, 8, 3)) == 64               # BASELINE configs[3]: 64 streams per GPU
    assert owner_of(130, 512, 8) == (2, 2)
    with pytest.raises(ValueError):
        shard_streams(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 7
        mine = shard_streams(n, world, rank)
        # one broadcast of the "weight blob"
        blob = torch.arange(1000, dtype=torch.uint8) if rank == 0 else torch.zeros(1000, dtype=torch.uint8)
        broadcast_weights(blob, 0)
        ok_blob = bool(torch.equal(blob, torch.arange(1000, dtype=torch.uint8)))
        # per-stream records computed locally (no per-frame collective), gathered for reporting
        local = torch.tensor([[float(s), float(s) * 2] for s in mine])
        allrec = gather_stream_records(local, n)
        slow = max_over_ranks(10.0 + rank)
        q.put((rank, list(mine), ok_blob, None if allrec is None else allrec.tolist(), slow))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_job():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, s0, b0, rec0, t0), (r1, s1, b1, rec1, t1) = res
    assert s0 == [0, 1, 2, 3] and s1 == [4, 5, 6]
    assert b0 and b1
    assert rec1 is None and rec0 == [[float(s), 2.0 * s] for s in range(7)]
    assert t0 == t1 == 11.0
