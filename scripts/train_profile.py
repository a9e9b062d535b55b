#!/usr/bin/env python
"""torch-profiler breakdown of one training step of scripts/train_bench.py's fusion stack (1 GPU)."""
import os, sys
import torch
from torch.profiler import ProfilerActivity, profile
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from train_bench import FusionStack
bf16 = "--bf16" in sys.argv
dev = "cuda"
torch.manual_seed(0)
dims, sizes = (256, 512, 1024), (80, 40, 20)
model = FusionStack(dims).to(dev)
opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9)
feats = [(torch.randn(16, d, s, s, device=dev), torch.randn(16, d, s, s, device=dev)) for d, s in zip(dims, sizes)]
def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16):
        loss = model(feats)
    loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.self_device_time_total)
tot = sum(e.self_device_time_total for e in rows)
print(f"total CUDA time per step {tot/3/1e3:.2f} ms")
for e in rows[:22]:
    print(f"{e.self_device_time_total/3/1e3:8.3f} ms {100*e.self_device_time_total/tot:5.1f}%  x{e.count//3:<3d} {e.key[:110]}")
