#!/usr/bin/env python
"""scripts/ffm_module_bench.py -- the non-transformer part of GPT1_fourier.forward (models/common.py:357-552) at the
reference's call-site shape, 2 x (B, 128, 160, 160): pooling -> pattern path -> (identity transformer) -> LayerNorm ->
token scatter -> bilinear upsample, forward and forward+backward.
`ours` = ffm.fourier_forward on a stand-in module; `stock` = the same op sequence with torch kernels only (torch.fft
Fourier split, 1x1 conv2d, closed-form separation loss, permute/contiguous, F.interpolate) -- i.e. what the reference
runs on the GPU, with its O(l^2) Python loss loop already replaced by the closed form.
Also times the four resampling kernels alone against the HBM roofline (bytes = one pass over the big map)."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffm_bench import sep_closed, torch_pattern  # noqa: E402
from ffm_glue_probe import PEAK, timed  # noqa: E402
from mmidet_b200 import ffm, ops  # noqa: E402


class StandIn(torch.nn.Module):
    def __init__(self, C):
        super().__init__()
        self.n_embd, self.vert_anchors, self.horz_anchors = C, 8, 8
        self.pos_emb = torch.nn.Parameter(torch.zeros(1, 128, C))
        self.trans_blocks = torch.nn.Sequential()
        self.ln_f = torch.nn.LayerNorm(C)
        self.drop = torch.nn.Dropout(0.1)
        self.avgpool = torch.nn.AdaptiveAvgPool2d((8, 8))
        self.conv1 = torch.nn.Conv2d(C, 8, 1, bias=False)
        self.conv2 = torch.nn.Conv2d(8, C, 1, bias=False)


def stock_forward(m, x):
    bs, c, h, w = x[0].shape
    tok, loss = torch_pattern(m.avgpool(x[0]), m.avgpool(x[1]), m.conv1, m.conv2, sep_closed)
    t = m.ln_f(m.trans_blocks(m.drop(m.pos_emb + tok)))
    t = t.view(bs, 2, 8, 8, c).permute(0, 1, 4, 2, 3)
    ro = F.interpolate(t[:, 0].contiguous(), size=[h, w], mode="bilinear")
    io = F.interpolate(t[:, 1].contiguous(), size=[h, w], mode="bilinear")
    return ro, io, loss


def main():
    torch.manual_seed(0)
    B, C, H, W = 16, 128, 160, 160
    for dt in (torch.float32, torch.bfloat16):
        m = StandIn(C).cuda().eval()
        vis = torch.randn(B, C, H, W, device="cuda", dtype=dt, requires_grad=True)
        ir = torch.randn(B, C, H, W, device="cuda", dtype=dt, requires_grad=True)
        g1, g2 = torch.randn_like(vis), torch.randn_like(ir)
        params = [vis, ir, m.conv1.weight, m.conv2.weight, m.pos_emb]

        def run(fwd, bwd):
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dt == torch.bfloat16):
                ro, io, _ = fwd(m, [vis, ir])
            if bwd:
                torch.autograd.grad([ro, io], params, [g1.to(ro.dtype), g2.to(io.dtype)])

        row = {"shape": [B, C, H, W], "dtype": str(dt)}
        for name, fwd in (("ours", ffm.fourier_forward), ("stock", stock_forward)):
            row[name + "_fwd_ms"] = round(timed(lambda: run(fwd, False)), 3)
            row[name + "_fwd_bwd_ms"] = round(timed(lambda: run(fwd, True)), 3)
        print(json.dumps(row), flush=True)
        big = vis.numel() * vis.element_size()
        small = torch.randn(B, C, 8, 8, device="cuda", dtype=dt)
        x = vis.detach()
        k = {"avgpool_fwd": lambda: ops._resample("mmi_avgpool_fwd", x, (B, C, 8, 8), True),
             "avgpool_bwd": lambda: ops._resample("mmi_avgpool_bwd", small, (B, C, H, W), False),
             "upsample_fwd": lambda: ops._resample("mmi_upsample_bilinear_fwd", small, (B, C, H, W), False),
             "upsample_bwd": lambda: ops._resample("mmi_upsample_bilinear_bwd", x, (B, C, 8, 8), True)}
        out = {"kernels": str(dt), "map_bytes": big}
        for name, fn in k.items():
            ms = timed(fn)
            out[name] = {"ms": round(ms, 4), "GBps": round(big / ms / 1e6, 1), "frac_of_peak": round(big / ms / 1e6 / PEAK, 3)}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
