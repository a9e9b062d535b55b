#!/usr/bin/env python
"""One forward (+ optionally backward) launch of the fused scan for ncu captures.
Usage: python scripts/prof_one.py [--B 16 --L 6400 --ED 512 --cfg 2 --bwd --randA --dtype f32]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmidet_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=16)
ap.add_argument("--L", type=int, default=6400)
ap.add_argument("--ED", type=int, default=512)
ap.add_argument("--cfg", type=int, default=0)
ap.add_argument("--bwd", action="store_true")
ap.add_argument("--randA", action="store_true")
ap.add_argument("--dtype", default="f32")
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
dt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[a.dtype]
B, L, ED, N = a.B, a.L, a.ED, 16
torch.manual_seed(0)
dev = "cuda"
x = torch.randn(B, L, ED, device=dev).to(dt)
delta = torch.nn.functional.softplus(torch.randn(B, L, ED, device=dev) - 3).to(dt)
z = torch.randn(B, L, ED, device=dev).to(dt)
Bm, Cm = torch.randn(2, B, L, N, device=dev).to(dt)
dout = torch.randn(B, L, ED, device=dev).to(dt)
D = torch.ones(ED, device=dev)
A = -torch.arange(1, N + 1, device=dev, dtype=torch.float32).repeat(ED, 1)
if a.randA:
    A = -torch.exp(torch.randn(ED, N, device=dev) * 0.7 + 0.5)
flags = a.cfg << 4
for _ in range(a.reps):
    out, _, chk, saved = ops.selscan_fwd_raw(x, delta, A, Bm, Cm, D, z=z, want_chk=True, flags=flags)
    if a.bwd:
        ops.selscan_bwd_raw(saved, chk, dout, flags=flags)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
