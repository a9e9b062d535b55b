#!/usr/bin/env python
"""Developer sweep (not the contract bench): CUDA-event timing of the scan kernels over lane mappings / A paths.
Usage: python scripts/devbench.py [--B 16 --L 6400 --ED 512 --dtype f32 --iters 10]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmidet_b200 import ops  # noqa: E402


def alg_bytes(B, L, ED, N, s, gate=True):
    fwd = B * L * ED * s * (4 if gate else 3) + B * L * N * s * 2 + (ED * N + ED) * 4
    bwd = B * L * ED * s * (7 if gate else 5) + B * L * N * s * 4 + 2 * (ED * N + ED) * 4
    return fwd, bwd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=16)
    ap.add_argument("--L", type=int, default=6400)
    ap.add_argument("--ED", type=int, default=512)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--cfgs", default="0,1,2,3")
    ap.add_argument("--nobwd", action="store_true")
    a = ap.parse_args()
    dt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[a.dtype]
    B, L, ED, N = a.B, a.L, a.ED, 16
    dev = "cuda"
    torch.manual_seed(0)
    x = torch.randn(B, L, ED, device=dev).to(dt)
    delta = torch.nn.functional.softplus(torch.randn(B, L, ED, device=dev) - 3).to(dt)
    z = torch.randn(B, L, ED, device=dev).to(dt)
    Bm, Cm = torch.randn(2, B, L, N, device=dev).to(dt)
    dout = torch.randn(B, L, ED, device=dev).to(dt)
    D = torch.ones(ED, device=dev)
    A_init = -torch.arange(1, N + 1, device=dev, dtype=torch.float32).repeat(ED, 1)
    A_rand = -torch.exp(torch.randn(ED, N, device=dev) * 0.7 + 0.5)
    fb, bb = alg_bytes(B, L, ED, N, x.element_size())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    print(f"B={B} L={L} ED={ED} dtype={a.dtype}  alg bytes fwd {fb/1e6:.1f} MB bwd {bb/1e6:.1f} MB")

    def timeit(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(a.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    for name, A in (("geomA", A_init), ("randA", A_rand)):
      for nseg in [0]:
        for lpc in [int(v) for v in a.cfgs.split(",")]:
            flags = lpc << 4
            t_inf = timeit(lambda: ops.selscan_fwd_raw(x, delta, A, Bm, Cm, D, z=z, flags=flags))
            t_f = timeit(lambda: ops.selscan_fwd_raw(x, delta, A, Bm, Cm, D, z=z, want_chk=True, flags=flags))
            line = (f"{name} cfg={lpc}: fwd(no chk) {t_inf:.3f} ms {fb/t_inf/1e6:.0f} GB/s | fwd(+chk) {t_f:.3f} ms "
                    f"{fb/t_f/1e6:.0f} GB/s")
            if not a.nobwd:
                _, _, chk, saved = ops.selscan_fwd_raw(x, delta, A, Bm, Cm, D, z=z, want_chk=True, flags=flags)
                t_b = timeit(lambda: ops.selscan_bwd_raw(saved, chk, dout, flags=flags))
                line += (f" | bwd {t_b:.3f} ms {bb/t_b/1e6:.0f} GB/s | fwd+bwd {(fb+bb)/(t_f+t_b)/1e6:.0f} GB/s "
                         f"= {(fb+bb)/(t_f+t_b)/1e6/6538*100:.1f}% of 6538")
            print(line, flush=True)


if __name__ == "__main__":
    main()
