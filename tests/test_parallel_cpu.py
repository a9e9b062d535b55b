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


def _detector_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    from torch.nn.parallel import DistributedDataParallel as DDP
    from mmidet_b200 import harness as H
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ref = H.import_reference()
        model = H.build_detector("s", "pytorch", seed=0, device="cpu").train()  # the reference's own blocks: runs on CPU
        hyp = H.scale_hyp(model, 6, 64)
        loss_fn = ref.loss.ComputeLoss(model)
        opt = H.make_optimizer(model, hyp, 2 * world)
        net = DDP(model, bucket_cap_mb=8, gradient_as_bucket_view=True, broadcast_buffers=False)
        imgs, targets = H.synthetic_batch(2, 64, seed=100 + rank, device="cpu")
        w0 = model.model[0].conv.conv.weight.detach().clone()
        f = imgs.float() / 255.0
        pred, comb = net(f[:, :3], f[:, 3:])
        loss, _ = loss_fn(pred, targets, comb.reshape(-1))
        (loss * world).backward()  # train.py:790-791
        g = model.model[0].conv.conv.weight.grad.detach().clone()
        opt.step()
        q.put((rank, float(loss), g.double().sum().item(), g.double().abs().sum().item(),
               float((model.model[0].conv.conv.weight.detach() - w0).abs().max())))
    finally:
        dist.destroy_process_group()


def test_detector_ddp_step_two_rank_gloo():
    """BASELINE configs[3] in miniature on CPU: the harness's DDP training step (unmodified reference Model / ComputeLoss,
    different synthetic batches per rank, loss * world_size, 8 MB buckets) on two gloo ranks: both ranks end up with the
    same all-reduced gradient and the optimizer moves the weights."""
    from mmidet_b200 import harness as H
    try:
        H.locate_reference()
    except RuntimeError:
        pytest.skip("no reference checkout")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_detector_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got[0][1] != got[1][1]  # different batches, different local losses
    assert abs(got[0][2] - got[1][2]) <= 1e-9 * max(1.0, abs(got[0][2])) and abs(got[0][3] - got[1][3]) <= 1e-9 * got[0][3]
    assert got[0][4] > 0 and got[1][4] > 0
