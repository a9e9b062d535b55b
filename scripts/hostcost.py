#!/usr/bin/env python
"""CPU-side cost of one operator call (python + ctypes + tensor-map encode + launch), measured with the GPU queue kept
short (small problem) so the host is what is timed."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmidet_b200 import ops
B, L, ED, N = 2, 64, 64, 16
dev = "cuda"
x = torch.randn(B, L, ED, device=dev); delta = torch.rand(B, L, ED, device=dev) * 0.1; z = torch.randn(B, L, ED, device=dev)
Bm, Cm = torch.randn(2, B, L, N, device=dev); dout = torch.randn(B, L, ED, device=dev)
A = -torch.arange(1, N + 1, device=dev, dtype=torch.float32).repeat(ED, 1); D = torch.ones(ED, device=dev)
for _ in range(20):
    out, _, chk, saved = ops.selscan_fwd_raw(x, delta, A, Bm, Cm, D, z=z, want_chk=True)
    ops.selscan_bwd_raw(saved, chk, dout)
torch.cuda.synchronize()
n = 200
t0 = time.perf_counter()
for _ in range(n):
    out, _, chk, saved = ops.selscan_fwd_raw(x, delta, A, Bm, Cm, D, z=z, want_chk=True)
torch.cuda.synchronize(); t1 = time.perf_counter()
for _ in range(n):
    ops.selscan_bwd_raw(saved, chk, dout)
torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host cost per call: fwd {1e6*(t1-t0)/n:.1f} us, bwd {1e6*(t2-t1)/n:.1f} us (includes tiny kernels)")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(100):
    ops.selscan_bwd_raw(saved, chk, dout)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
