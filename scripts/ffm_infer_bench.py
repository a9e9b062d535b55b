#!/usr/bin/env python
"""scripts/ffm_infer_bench.py -- batch-1 latency of the FFM forward outside its transformer (GPT1_fourier call site of a
640 px two-stream YOLOv5: 2 x (1, 128, 160, 160), fp16): stock torch op sequence vs ffm.fourier_forward eager vs the same
replayed as one CUDA graph (graphs.Graphed)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffm_module_bench import StandIn, stock_forward  # noqa: E402
from mmidet_b200 import ffm  # noqa: E402
from mmidet_b200.graphs import Graphed  # noqa: E402


class Wrap(torch.nn.Module):
    def __init__(self, inner):
        super().__init__()
        self.inner = inner

    def forward(self, x):
        return ffm.fourier_forward(self.inner, x)


def lat(fn, iters=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / iters


def main():
    torch.manual_seed(0)
    C = 128
    m = StandIn(C).cuda().half().eval()
    vis = torch.randn(1, C, 160, 160, device="cuda").half()
    ir = torch.randn(1, C, 160, 160, device="cuda").half()
    fast = Graphed(Wrap(m))
    with torch.no_grad():
        row = {"shape": [1, C, 160, 160], "dtype": "float16",
               "stock_ms": round(lat(lambda: stock_forward(m, [vis, ir])), 4),
               "ours_eager_ms": round(lat(lambda: ffm.fourier_forward(m, [vis, ir])), 4),
               "ours_graph_ms": round(lat(lambda: fast([vis, ir])), 4)}
    print(json.dumps(row))


if __name__ == "__main__":
    main()
