"""pscan -- drop-in for `models.pscan.pscan` (= PScan.apply, models/pscan.py:226).

Same call: pscan(A_in, X_in) with A_in, X_in (B, L, D, N) -> H (B, L, D, N), H[t] = A[t] * H[t-1] + X[t];
backward returns (gradA, gradX) exactly as PScan.backward (models/pscan.py:189-224).  Inputs are not modified
(the reference clones, :167-174).  Runs on the CUDA kernels of csrc/pscan.cu through the C ABI: float64 inputs are scanned
in float64, every other dtype in fp32 (converted on entry, result cast back)."""
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
        if A_in.shape != X_in.shape or X_in.dim() != 4:
            raise ValueError(f"pscan: A_in and X_in must both be (B, L, D, N); got {tuple(A_in.shape)} and {tuple(X_in.shape)}")
        lib = _lib.load()
        B, L, D, N = X_in.shape
        # the reference is dtype-generic (models/pscan.py:37-92): float64 inputs are scanned in float64, everything else in fp32
        f64 = X_in.dtype == torch.float64 or A_in.dtype == torch.float64
        wd = torch.float64 if f64 else torch.float32
        A = A_in.detach().to(wd).contiguous()
        X = X_in.detach().to(wd).contiguous()
        H = torch.empty_like(X)
        ws = _ws(lib, B, L, D, N, X.device)
        fn = lib.mmi_pscan_fwd_f64 if f64 else lib.mmi_pscan_fwd
        _lib.check(fn(_ops._ptr(A), _ops._ptr(X), _ops._ptr(H), _ops._ptr(ws), B, L, D, N, _ops._stream(X)), "mmi_pscan_fwd")
        _ops.launches += 2
        ctx.save_for_backward(A, H)
        ctx.dtypes = (A_in.dtype, X_in.dtype)
        ctx.f64 = f64
        return H.to(X_in.dtype)

    @staticmethod
    def backward(ctx, grad_output_in):
        lib = _lib.load()
        A, H = ctx.saved_tensors
        B, L, D, N = H.shape
        g = grad_output_in.to(H.dtype).contiguous()
        gA, gX = torch.empty_like(H), torch.empty_like(H)
        ws = _ws(lib, B, L, D, N, H.device)
        fn = lib.mmi_pscan_bwd_f64 if ctx.f64 else lib.mmi_pscan_bwd
        _lib.check(fn(_ops._ptr(A), _ops._ptr(H), _ops._ptr(g), _ops._ptr(gA), _ops._ptr(gX), _ops._ptr(ws), B, L, D, N,
                      _ops._stream(H)), "mmi_pscan_bwd")
        _ops.launches += 2
        return gA.to(ctx.dtypes[0]), gX.to(ctx.dtypes[1])


pscan = PScan.apply
