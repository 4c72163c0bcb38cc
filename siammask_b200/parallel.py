"""Multi-GPU host logic of the hot path: tracker streams are independent, so they are sharded
across ranks in contiguous blocks; the only collective of the whole job is ONE broadcast of the packed weight
arena at init (NCCL over NVLink on GPUs; the same code runs on gloo/CPU tensors in the tests).  Results stay on
the rank that owns the stream; `gather_stream_records` collects small per-stream records on rank 0 for reporting.
Reference precedent: process-per-GPU fan-out without communication (experiments/siammask_sharp/test_all.sh:66-70).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_streams(num_streams: int, world: int, rank: int) -> range:
    """Contiguous block of stream ids owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world) or num_streams < 0:
        raise ValueError("bad rank/world/num_streams")
    base, rem = divmod(num_streams, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def owner_of(stream_id: int, num_streams: int, world: int) -> tuple[int, int]:
    """(rank, local slot) of a global stream id under `shard_streams`."""
    for r in range(world):
        rg = shard_streams(num_streams, world, r)
        if stream_id in rg:
            return r, stream_id - rg.start
    raise IndexError(stream_id)


def broadcast_weights(blob: torch.Tensor, src: int = 0) -> torch.Tensor:
    """The one collective of the job: rank `src` holds the packed weights, everybody else receives them in place."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(blob, src=src)
    return blob


def max_over_ranks(value: float, device=None) -> float:
    """Timing contract of bench.py: a step takes as long as the slowest rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_stream_records(local: torch.Tensor, num_streams: int) -> torch.Tensor | None:
    """local: [n_local, K] records of this rank's streams (in slot order) -> [num_streams, K] on rank 0, else None."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [len(shard_streams(num_streams, world, r)) for r in range(world)]
    pad = max(sizes)
    buf = torch.zeros(pad, local.shape[1], dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0)
    if rank != 0:
        return None
    return torch.cat([o[:n] for o, n in zip(out, sizes)], 0)
