"""Multi-GPU plumbing of the hot path: the scan couples only the timesteps of one (batch, channel) row, so
the path shards over the BATCH with no data-path collective -- one process per GPU, each rank scans its own image pairs.
The only exchange in training is the gradient all-reduce of the replicated parameters (reference call site: the DDP
wrapper, train.py:684), done by NCCL over NVLink; timings are combined as the max over ranks.  Everything here is
backend-agnostic torch.distributed so the logic is covered on CPU with gloo (tests/test_parallel_cpu.py)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_rank_world():
    """(rank, local_rank, world) from the torchrun environment; (0, 0, 1) when launched plainly."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard(n_units: int, rank: int, world: int):
    """[start, stop) of the contiguous block of units (image pairs) owned by `rank`; blocks differ by at most one unit."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_units, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def max_over_ranks(values, device="cpu", group=None):
    """Element-wise max over ranks of a list of floats (device times are reported as the slowest rank's)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return [float(v) for v in t]


def allreduce_mean_grads(params, bucket_bytes: int = 64 << 20, group=None):
    """Average .grad of replicated parameters over ranks in flat buckets (sized for launch latency, not link count:
    NVSwitch gives every peer full bandwidth).  Equivalent to what DistributedDataParallel does for the fusion blocks;
    used by harnesses that drive the kernels without the DDP wrapper."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    grads = [p.grad for p in params if p.grad is not None]
    buckets, cur, size = [], [], 0
    for g in grads:
        n = g.numel() * g.element_size()
        if cur and (size + n > bucket_bytes or g.dtype != cur[0].dtype):
            buckets.append(cur)
            cur, size = [], 0
        cur.append(g)
        size += n
    if cur:
        buckets.append(cur)
    for b in buckets:
        flat = torch.cat([g.reshape(-1) for g in b])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        off = 0
        for g in b:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
    return len(buckets)
