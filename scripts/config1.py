#!/usr/bin/env python
"""BASELINE configs[0]: MambaBlock / selective scan standalone forward at B=2, L=6400 (80x80 tokens), d_inner=256, d_state=16,
fp32 -- GPU (this library) next to the CPU oracle port on the box's host cores."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmidet_b200 import ops
from mmidet_b200.mamba import MambaBlock, MambaConfig
from oracle import oracle as O

def gpu_time(fn, it=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it

torch.manual_seed(0)
B, L, D = 2, 6400, 128
blk = MambaBlock(MambaConfig(d_model=D, n_layers=1)).cuda()
x = torch.randn(B, L, D, device="cuda")
with torch.no_grad():
    t_blk = gpu_time(lambda: blk(x))
xr = x.clone().requires_grad_(True)
g = torch.randn(B, L, D, device="cuda")
t_blk_fb = gpu_time(lambda: blk(xr).backward(g))
ED, N = 2 * D, 16
xs = torch.randn(B, L, ED, device="cuda"); delta = torch.nn.functional.softplus(torch.randn(B, L, ED, device="cuda") - 3)
Bm, Cm = torch.randn(2, B, L, N, device="cuda")
A = -torch.arange(1, N + 1, device="cuda", dtype=torch.float32).repeat(ED, 1); Dv = torch.ones(ED, device="cuda")
t_scan = gpu_time(lambda: ops.selscan_fwd_raw(xs, delta, A, Bm, Cm, Dv))
_, _, chk, saved = ops.selscan_fwd_raw(xs, delta, A, Bm, Cm, Dv, want_chk=True)
dout = torch.randn(B, L, ED, device="cuda")
t_scan_b = gpu_time(lambda: ops.selscan_bwd_raw(saved, chk, dout))
cores = os.cpu_count(); ol = O.lib(); ol.oracle_set_threads(cores)
a = [t.cpu().numpy() for t in (xs, delta, A, Bm, Cm, Dv, dout)]
O.selective_scan_fwd(*a[:6]); t0 = time.perf_counter(); O.selective_scan_fwd(*a[:6]); t_cpu_f = time.perf_counter() - t0
t0 = time.perf_counter(); O.selective_scan_bwd(*a[:6], a[6]); t_cpu_b = time.perf_counter() - t0
print(f"configs[0] B={B} L={L} d_inner={ED} fp32: MambaBlock fwd {t_blk:.3f} ms, fwd+bwd {t_blk_fb:.3f} ms (GPU, whole block); "
      f"scan fwd {t_scan*1e3:.1f} us, scan bwd {t_scan_b*1e3:.1f} us (GPU) vs CPU oracle port on {cores} threads: fwd {t_cpu_f*1e3:.1f} ms, bwd {t_cpu_b*1e3:.1f} ms")
