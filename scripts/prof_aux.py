#!/usr/bin/env python
"""scripts/prof_aux.py -- launches the RMSNorm / causal-conv backward kernels at the P3 training shape (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmidet_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
P, DT, ST = ops._ptr, ops._DT, ops._stream
dt = torch.float32 if "--fp32" in sys.argv else torch.bfloat16
Bt, C, Lt, ED = 16, 256, 12800, 512
x2 = torch.randn(Bt * Lt, C, device="cuda", dtype=dt)
g2 = torch.randn_like(x2)
dx2, dw2, w32 = torch.empty_like(x2), torch.empty(C, device="cuda"), torch.ones(C, device="cuda")
xc = torch.randn(Bt, Lt, ED, device="cuda", dtype=dt)
gy, dxc = torch.randn_like(xc), torch.empty_like(xc)
wc, bc = torch.randn(ED, 4, device="cuda"), torch.randn(ED, device="cuda")
dwc, dbc = torch.empty(ED, 4, device="cuda"), torch.empty(ED, device="cuda")
for _ in range(3):
    lib.mmi_rmsnorm_bwd(P(x2), P(w32), P(g2), P(dx2), P(dw2), x2.shape[0], C, x2.stride(0), g2.stride(0), dx2.stride(0), 1e-5, DT[dt], -1, ST(x2))
    lib.mmi_causal_conv1d_bwd(P(xc), P(wc), P(bc), P(gy), P(dxc), P(dwc), P(dbc), Bt, Lt, ED, 4, xc.stride(1), gy.stride(1), dxc.stride(1),
                              DT[dt], 1, ST(xc))
torch.cuda.synchronize()
