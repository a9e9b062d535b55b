"""pscan -- drop-in for `models.pscan.pscan` (= PScan.apply, models/pscan.py:226).

Same call: pscan(A_in, X_in) with A_in, X_in (B, L, D, N) -> H (B, L, D, N), H[t] = A[t] * H[t-1] + X[t];
backward returns (gradA, gradX) exactly as PScan.backward (models/pscan.py:189-224).  Inputs are not modified
(the reference clones, :167-174).  Runs on the CUDA kernels of csrc/pscan.cu through the C ABI; fp32 compute
(other dtypes are converted on entry and the result cast back)."""
from __future__ import annotations

import ctypes
import math

import torch

from . import _lib
from . import ops as _ops


def npo2(length: int) -> int:
    """models/pscan.py:13-18 (kept for API compatibility; the CUDA scan needs no padding)."""
    return 2 ** math.ceil(math.log2(length))


def _ws(lib, B, L, D, N, device):
    return torch.empty(max(16, lib.mmi_pscan_ws_bytes(B, L, D, N)), dtype=torch.uint8, device=device)


class PScan(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A_in, X_in):
        if not A_in.is_cuda:
            raise RuntimeError("mmidet_b200.pscan: CUDA tensors required (no CPU path)")
        lib = _lib.load()
        B, L, D, N = X_in.shape
        A = A_in.detach().float().contiguous()
        X = X_in.detach().float().contiguous()
        H = torch.empty_like(X)
        ws = _ws(lib, B, L, D, N, X.device)
        _lib.check(lib.mmi_pscan_fwd(_ops._ptr(A), _ops._ptr(X), _ops._ptr(H), _ops._ptr(ws), B, L, D, N,
                                     _ops._stream(X)), "mmi_pscan_fwd")
        _ops.launches += 2
        ctx.save_for_backward(A, H)
        ctx.dtypes = (A_in.dtype, X_in.dtype)
        return H.to(X_in.dtype)

    @staticmethod
    def backward(ctx, grad_output_in):
        lib = _lib.load()
        A, H = ctx.saved_tensors
        B, L, D, N = H.shape
        g = grad_output_in.float().contiguous()
        gA, gX = torch.empty_like(H), torch.empty_like(H)
        ws = _ws(lib, B, L, D, N, H.device)
        _lib.check(lib.mmi_pscan_bwd(_ops._ptr(A), _ops._ptr(H), _ops._ptr(g), _ops._ptr(gA), _ops._ptr(gX),
                                     _ops._ptr(ws), B, L, D, N, _ops._stream(H)), "mmi_pscan_bwd")
        _ops.launches += 2
        return gA.to(ctx.dtypes[0]), gX.to(ctx.dtypes[1])


pscan = PScan.apply
