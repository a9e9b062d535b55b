#!/usr/bin/env python
"""pscan parity API (materialised (B, L, D, N) tensors) timing: python scripts/pscan_bench.py [--B 2 --L 6400 --D 256]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmidet_b200.pscan import pscan  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=2)
ap.add_argument("--L", type=int, default=6400)
ap.add_argument("--D", type=int, default=256)
a = ap.parse_args()
N = 16
A = (torch.rand(a.B, a.L, a.D, N, device="cuda") * 0.3 + 0.7).requires_grad_(True)
X = torch.randn(a.B, a.L, a.D, N, device="cuda", requires_grad=True)
g = torch.randn(a.B, a.L, a.D, N, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
nbytes = A.numel() * 4


def timeit(fn, n=7):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


H = pscan(A, X)
tf = timeit(lambda: pscan(A, X))
tb = timeit(lambda: torch.autograd.grad(H, (A, X), g, retain_graph=True))
print(f"pscan B={a.B} L={a.L} D={a.D} N={N} fp32 ({nbytes/1e6:.0f} MB per tensor): forward {tf:.3f} ms = {3*nbytes/tf/1e6:.0f} GB/s "
      f"over 3 passes = {100*3*nbytes/tf/1e6/6538:.1f} % of 6538; backward {tb:.3f} ms = {5*nbytes/tb/1e6:.0f} GB/s over 5 passes "
      f"= {100*5*nbytes/tb/1e6/6538:.1f} %")
