#!/usr/bin/env python
"""Where the time goes inside one MambaBlock fwd+bwd on the GPU (torch profiler, CUDA time per kernel family)."""
import os, sys
import torch
from torch.profiler import ProfilerActivity, profile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmidet_b200.mamba import MambaBlock, MambaConfig
B, L, D = int(os.environ.get("B", 16)), 6400, 256
torch.manual_seed(0)
blk = MambaBlock(MambaConfig(d_model=D, n_layers=1)).cuda()
x = torch.randn(B, L, D, device="cuda", requires_grad=True)
g = torch.randn(B, L, D, device="cuda")
for _ in range(3):
    blk(x).backward(g)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5):
        blk(x).backward(g)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
