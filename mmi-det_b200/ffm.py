"""Fusion Focus Module Fourier step -- drop-ins for the helpers of models/common.py.

extract_frequency2(image) -> (low, high): same contract as models/common.py:37-69 (both outputs torch.float16,
real, shape of `image`), including the negative-slice wrap of the reference's masks (:44-56).
separation_loss(M): models/common.py:128-139 (`Seperation_loss`) in closed form.
Both run on csrc/ffm.cu through the C ABI; CUDA tensors only."""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from . import ops as _ops


def kept_range(H: int, W: int):
    """[r0, r1) x [c0, c1): block of the fftshift-ed spectrum the reference's low-pass keeps / high-pass zeroes."""
    lib = _lib.load()
    v = [ctypes.c_int() for _ in range(4)]
    lib.mmi_ffm_kept_range(H, W, *[ctypes.byref(i) for i in v])
    return tuple(i.value for i in v)


def extract_frequency2(image: torch.Tensor, with_product: bool = False):
    """-> (low, high) float16; with_product=True also returns high * image in fp32 (models/common.py:440-441)."""
    if not image.is_cuda:
        raise RuntimeError("mmidet_b200.extract_frequency2: CUDA tensor required (no CPU path)")
    lib = _lib.load()
    Bsz, C, H, W = image.shape
    if image.dtype not in _ops._DT:
        image = image.float()
    img = image.contiguous()
    low = torch.empty((Bsz, C, H, W), dtype=torch.float16, device=img.device)
    high = torch.empty_like(low)
    prod = torch.empty((Bsz, C, H, W), dtype=torch.float32, device=img.device) if with_product else None
    _lib.check(lib.mmi_ffm_extract(_ops._ptr(img), _ops._ptr(low), _ops._ptr(high), _ops._ptr(prod), Bsz * C, H, W,
                                   _ops._DT[img.dtype], _ops._stream(img)), "mmi_ffm_extract")
    _ops.launches += 1
    return (low, high, prod) if with_product else (low, high)


def fourier_transform(image: torch.Tensor) -> torch.Tensor:
    """models/common.py:25-32: fftshift(fftn(image)) -- plain cuFFT through torch.fft (library call, not on the
    reference's forward path: its only caller `extract_frequency` is dead code)."""
    return torch.fft.fftshift(torch.fft.fftn(image, dim=(-2, -1)), dim=(-2, -1))


def separation_loss(M: torch.Tensor) -> torch.Tensor:
    """models/common.py:128-139 for M (l, K): sum_{i<j} M_i . M_j / (l (l - 1)), as a 0-d fp32 tensor."""
    if not M.is_cuda:
        raise RuntimeError("mmidet_b200.separation_loss: CUDA tensor required (no CPU path)")
    lib = _lib.load()
    Mc = M.detach().float().contiguous()
    out = torch.empty(1, dtype=torch.float32, device=M.device)
    _lib.check(lib.mmi_separation_loss(_ops._ptr(Mc), _ops._ptr(out), Mc.shape[0], Mc.shape[1], _ops._stream(Mc)),
               "mmi_separation_loss")
    _ops.launches += 1
    return out[0]


Seperation_loss = separation_loss  # the reference's spelling (models/common.py:128)
