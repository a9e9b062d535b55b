#!/usr/bin/env python
"""Inference latency of the fusion blocks at the four call sites of a 640 px two-stream YOLOv5s (BASELINE configs[1] shapes:
d_model 64/128/256/512 at 160x160/80x80/40x40/20x20, batch 1), no_grad, fp32 and fp16 (detect_twostream.py runs model.half())."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmidet_b200.graphs import Graphed
from mmidet_b200.mamba import MambaFusion

def gpu_time(fn, it=50):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it

torch.manual_seed(0)
B = int(os.environ.get("B", 1))
for dt in (torch.float32, torch.float16):
    tot = totg = 0.0
    for d, s in ((64, 160), (128, 80), (256, 40), (512, 20)):
        m = MambaFusion(d, n_layer=1).cuda().to(dt).eval()
        x = [torch.randn(B, d, s, s, device="cuda", dtype=dt), torch.randn(B, d, s, s, device="cuda", dtype=dt)]
        with torch.no_grad():
            t = gpu_time(lambda: m(x))
        g = Graphed(m)
        tg = gpu_time(lambda: g(x))
        tot += t
        totg += tg
        print(f"{str(dt)[6:]:8s} B={B} d_model={d:4d} {s}x{s} (L={2*s*s:6d} tokens, d_inner={2*d:4d}): fusion forward {t*1e3:7.1f} us eager, "
              f"{tg*1e3:7.1f} us as a CUDA graph")
    print(f"{str(dt)[6:]:8s} four sites together: {tot:.3f} ms eager, {totg:.3f} ms graphed, per image pair")
