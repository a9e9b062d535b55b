#!/usr/bin/env python
"""Data-parallel training step of the fusion blocks: batch sharded over the ranks, replicated MambaFusion
blocks at the three detector scales (P3/P4/P5 of a 640 px two-stream YOLOv5l: L = 6400/1600/400 tokens per modality,
d_model = 256/512/1024), gradient all-reduce by DistributedDataParallel over NCCL.  The backbone convolutions are not part
of this repository's path, so synthetic feature maps stand in for them; what is measured is fusion forward + backward +
all-reduce + SGD step, in image pairs per second.

    python scripts/train_bench.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/train_bench.py
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmidet_b200.mamba import MambaFusion  # noqa: E402
from mmidet_b200.parallel import env_rank_world, max_over_ranks  # noqa: E402


class FusionStack(nn.Module):
    def __init__(self, dims):
        super().__init__()
        self.f = nn.ModuleList([MambaFusion(d, n_layer=1) for d in dims])

    def forward(self, feats):
        loss = 0.0
        for m, (rgb, ir) in zip(self.f, feats):
            a, b = m([rgb, ir])
            loss = loss + (a.float() ** 2).mean() + (b.float() ** 2).mean()
        return loss


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16, help="image pairs per GPU")
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--bf16", action="store_true", help="autocast the projections to bf16 (scan state stays fp32)")
    a = ap.parse_args()
    rank, local, world = env_rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)  # identical replicas
    dims, sizes = (256, 512, 1024), (80, 40, 20)
    model = FusionStack(dims).to(dev)
    ddp = nn.parallel.DistributedDataParallel(model, device_ids=[local], bucket_cap_mb=64) if world > 1 else model
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9)
    g = torch.Generator(device=dev).manual_seed(100 + rank)  # each rank owns its own shard of the global batch
    feats = [(torch.randn(a.batch, d, s, s, device=dev, generator=g), torch.randn(a.batch, d, s, s, device=dev, generator=g))
             for d, s in zip(dims, sizes)]

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=a.bf16):
            loss = ddp(feats)
        loss.backward()
        opt.step()
        return loss

    for _ in range(a.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    (ms,) = max_over_ranks([e0.elapsed_time(e1) / a.steps], device=dev)
    nparam = sum(p.numel() for p in model.parameters())
    if rank == 0:
        print(json.dumps({"metric": "fusion_train_pairs_per_s", "value": round(world * a.batch / (ms * 1e-3), 1), "n_gpus": world,
                          "ms_per_step": round(ms, 3), "pairs_per_gpu": a.batch, "params": nparam,
                          "allreduce_bytes_per_step": 4 * nparam if world > 1 else 0, "autocast_bf16": a.bf16,
                          "loss": round(float(loss), 6), "scales": list(zip(dims, sizes))}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
