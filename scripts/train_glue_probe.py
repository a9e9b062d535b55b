#!/usr/bin/env python
"""which stock-torch ops (by shape) remain in one bf16 training step of the fusion stack -- CPU-side op names with device time."""
import os, sys
import torch
from torch.profiler import ProfilerActivity, profile
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from train_bench import FusionStack
dev = "cuda"
torch.manual_seed(0)
dims, sizes = (256, 512, 1024), (80, 40, 20)
model = FusionStack(dims).to(dev)
opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9)
feats = [(torch.randn(16, d, s, s, device=dev), torch.randn(16, d, s, s, device=dev)) for d, s in zip(dims, sizes)]
def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = model(feats)
    loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True) if e.self_device_time_total > 30]
rows.sort(key=lambda e: -e.self_device_time_total)
for e in rows[:40]:
    print(f"{e.self_device_time_total/1e3:7.3f} ms x{e.count:<3d} {e.key[:40]:40s} {str(e.input_shapes)[:120]}")
