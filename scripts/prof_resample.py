#!/usr/bin/env python
"""scripts/prof_resample.py -- launches the four resampling directions at the FFM call-site shape (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmidet_b200 import ops  # noqa: E402

B, C, H, W = 16, 128, 160, 160
dt = torch.bfloat16 if "--bf16" in sys.argv else torch.float32
x = torch.randn(B, C, H, W, device="cuda", dtype=dt)
s = torch.randn(B, C, 8, 8, device="cuda", dtype=dt)
for _ in range(3):
    ops._resample("mmi_avgpool_fwd", x, (B, C, 8, 8), True)
    ops._resample("mmi_upsample_bilinear_bwd", x, (B, C, 8, 8), True)
    ops._resample("mmi_upsample_bilinear_fwd", s, (B, C, H, W), False)
    ops._resample("mmi_avgpool_bwd", s, (B, C, H, W), False)
torch.cuda.synchronize()
