"""world_size-2 gloo tests (CPU) of the batch-sharding / timing / gradient-averaging plumbing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from mmidet_b200 import parallel as P
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert P.env_rank_world() == (rank, rank, world)
        # 1. batch sharding: 7 image pairs over 2 ranks, replicated weights, gradient mean == single-process gradient
        torch.manual_seed(0)
        X, Y = torch.randn(7, 5), torch.randn(7, 3)
        ref = torch.nn.Linear(5, 3)
        model = torch.nn.Linear(5, 3)
        model.load_state_dict(ref.state_dict())
        a, b = P.shard(7, rank, world)
        # sum-of-squares loss scaled so that the MEAN over ranks of local grads equals the global gradient
        loss = ((model(X[a:b]) - Y[a:b]) ** 2).sum() * world / 7
        loss.backward()
        nb = P.allreduce_mean_grads(list(model.parameters()), bucket_bytes=32)  # tiny buckets: exercises bucketing
        ((ref(X) - Y) ** 2).sum().div(7).backward()
        for p, r in zip(model.parameters(), ref.parameters()):
            assert torch.allclose(p.grad, r.grad, atol=1e-6), (rank, (p.grad - r.grad).abs().max())
        # 2. max-over-ranks timing
        t = P.max_over_ranks([1.0 + rank, 5.0 - rank])
        assert t == [float(world), 5.0]
        q.put((rank, a, b, nb))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert [(g[1], g[2]) for g in got] == [(0, 4), (4, 7)]
    assert all(g[3] >= 2 for g in got)


@pytest.mark.parametrize("n,world", [(16, 8), (7, 2), (3, 4), (0, 2), (128, 3)])
def test_shard_partitions_exactly(n, world):
    from mmidet_b200.parallel import shard
    spans = [shard(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1
