#!/usr/bin/env python
"""scripts/ffm_glue_probe.py -- how close the stock torch kernels either side of the FFM pattern path run to the HBM
roofline at the reference's call-site shape (GPT1_fourier input 2 x (B, 128, 160, 160), common.py:396-397, :540-543):
AdaptiveAvgPool2d(8, 8) and bilinear upsample back to (160, 160), forward and backward."""
import json
import sys

import torch
import torch.nn.functional as F

PEAK = 6538.0


def timed(fn, iters=20, warm=5):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for i in range(warm + iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        if i >= warm:
            ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    B, C, H, W = 16, 128, 160, 160
    for dt in (torch.float32, torch.bfloat16):
        s = torch.empty(0, dtype=dt).element_size()
        x = torch.randn(B, C, H, W, device="cuda", dtype=dt, requires_grad=True)
        big = B * C * H * W * s
        p = F.adaptive_avg_pool2d(x, (8, 8))
        gp = torch.randn_like(p)
        small = torch.randn(B, C, 8, 8, device="cuda", dtype=dt, requires_grad=True)
        up = F.interpolate(small, size=[H, W], mode="bilinear")
        gu = torch.randn_like(up)
        rows = {
            "pool_fwd": timed(lambda: F.adaptive_avg_pool2d(x, (8, 8))),
            "pool_bwd": timed(lambda: torch.autograd.grad(p, x, gp, retain_graph=True)),
            "upsample_fwd": timed(lambda: F.interpolate(small, size=[H, W], mode="bilinear")),
            "upsample_bwd": timed(lambda: torch.autograd.grad(up, small, gu, retain_graph=True)),
        }
        out = {"dtype": str(dt), "map_bytes": big}
        for k, ms in rows.items():
            out[k] = {"ms": round(ms, 4), "GBps": round(big / ms / 1e6, 1), "frac_of_peak": round(big / ms / 1e6 / PEAK, 3)}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
