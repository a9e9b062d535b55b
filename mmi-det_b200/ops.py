"""Operator layer: torch.autograd.Function wrappers over the C ABI (include/mmidet_b200.h).

selective_scan(x, delta, A, B, C, D, z=None) mirrors MambaBlock.selective_scan of the reference
(models/mamba.py:212-233) -- same argument order, shapes and meaning -- with the SiLU gate of
MambaBlock.forward (models/mamba.py:184-186) optionally fused through `z`.
PyTorch is used for device memory, streams and autograd plumbing only; there is no eager fallback."""
from __future__ import annotations

import ctypes

import torch

from . import _lib

_DT = {torch.float32: _lib.MMI_F32, torch.bfloat16: _lib.MMI_BF16, torch.float16: _lib.MMI_F16}

# launch counter: number of kernels of OURS enqueued through this module (bench.py reports it as gpu_launches)
launches = 0


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _require_cuda(t, who):
    if not t.is_cuda:
        raise RuntimeError(f"{who}: expected a CUDA tensor; mmidet_b200 has no CPU path (got device {t.device})")


def _rows(t: torch.Tensor, esz_mult: int = 16) -> torch.Tensor:
    """Return `t` (B, L, E) in a layout the kernels can address as rows of pitch `ld` without copying when
    possible: unit channel stride, batch stride == L * row stride, 16-byte aligned rows."""
    B, L, E = t.shape
    ok = t.stride(2) == 1 and (B == 1 or t.stride(0) == L * t.stride(1)) and t.stride(1) >= E \
        and (t.stride(1) * t.element_size()) % 16 == 0 and t.data_ptr() % 16 == 0
    return t if ok else t.contiguous()


def selscan_chunk() -> int:
    return _lib.load().mmi_selscan_chunk()


def selscan_fwd_raw(x, delta, A, Bm, Cm, D, z=None, h0=None, want_state=False, want_chk=False, flags=0):
    """Direct call of mmi_selscan_fwd. Returns (out, hT or None, chk or None)."""
    global launches
    lib = _lib.load()
    _require_cuda(x, "selective_scan")
    if x.dim() != 3 or A.dim() != 2:
        raise ValueError(f"selective_scan: x must be (B, L, ED) and A (ED, N); got {tuple(x.shape)} and {tuple(A.shape)}")
    Bsz, L, ED = x.shape
    N = A.shape[1]
    # the C ABI takes raw pointers: every shape is checked here (the reference would raise a broadcasting error instead)
    want = {"delta": (delta, (Bsz, L, ED)), "z": (z, (Bsz, L, ED)), "A": (A, (ED, N)), "B": (Bm, (Bsz, L, N)),
            "C": (Cm, (Bsz, L, N)), "D": (D, (ED,)), "h0": (h0, (Bsz, ED, N))}
    for name, (t, shp) in want.items():
        if t is not None and tuple(t.shape) != shp:
            raise ValueError(f"selective_scan: {name} has shape {tuple(t.shape)}, expected {shp} for x {tuple(x.shape)}")
        if t is not None and t.device != x.device:
            raise RuntimeError(f"selective_scan: {name} is on {t.device}, x on {x.device}")
    if x.dtype not in _DT:
        raise RuntimeError(f"selective_scan: unsupported dtype {x.dtype} (float32 / bfloat16 / float16)")
    dt = x.dtype
    x, delta = _rows(x), _rows(delta.to(dt))
    z = None if z is None else _rows(z.to(dt))
    Bm, Cm = Bm.to(dt).contiguous(), Cm.to(dt).contiguous()
    A, D = A.float().contiguous(), D.float().contiguous()
    out = torch.empty((Bsz, L, ED), dtype=dt, device=x.device)
    hT = torch.empty((Bsz, ED, N), dtype=torch.float32, device=x.device) if want_state else None
    chunk = lib.mmi_selscan_chunk()
    chk = torch.empty((Bsz, (L + chunk - 1) // chunk, ED, N), dtype=torch.float32, device=x.device) if want_chk else None
    h0 = None if h0 is None else h0.float().contiguous()
    ws = torch.empty(lib.mmi_selscan_fwd_ws_bytes(Bsz, L, ED, N), dtype=torch.uint8, device=x.device)  # L-split summaries
    _lib.check(lib.mmi_selscan_fwd(_ptr(x), _ptr(delta), _ptr(z), _ptr(A), _ptr(Bm), _ptr(Cm), _ptr(D), _ptr(h0),
                                   _ptr(out), _ptr(hT), _ptr(chk), _ptr(ws), Bsz, L, ED, N, x.stride(1), delta.stride(1),
                                   z.stride(1) if z is not None else 0, out.stride(1), chunk, _DT[dt], flags,
                                   _stream(x)), "mmi_selscan_fwd")
    launches += 1
    return out, hT, chk, (x, delta, z, A, Bm, Cm, D)


def selscan_bwd_raw(saved, chk, dout, flags=0):
    """Direct call of mmi_selscan_bwd. `saved` is the normalised input tuple returned by selscan_fwd_raw."""
    global launches
    lib = _lib.load()
    x, delta, z, A, Bm, Cm, D = saved
    Bsz, L, ED = x.shape
    N = A.shape[1]
    dt = x.dtype
    if tuple(dout.shape) != (Bsz, L, ED):
        raise ValueError(f"selective_scan backward: dout has shape {tuple(dout.shape)}, expected {(Bsz, L, ED)}")
    nchk = (L + lib.mmi_selscan_chunk() - 1) // lib.mmi_selscan_chunk()
    if chk is None or tuple(chk.shape) != (Bsz, nchk, ED, N):
        raise ValueError("selective_scan backward: checkpoints of the matching forward call are required")
    dout = _rows(dout.to(dt))
    dx, dd = torch.empty((Bsz, L, ED), dtype=dt, device=x.device), torch.empty((Bsz, L, ED), dtype=dt, device=x.device)
    dz = torch.empty((Bsz, L, ED), dtype=dt, device=x.device) if z is not None else None
    dA, dD = torch.empty_like(A), torch.empty_like(D)
    dB, dC = torch.empty_like(Bm), torch.empty_like(Cm)
    ws = torch.empty(lib.mmi_selscan_bwd_ws_bytes(Bsz, L, ED, N), dtype=torch.uint8, device=x.device)
    _lib.check(lib.mmi_selscan_bwd(_ptr(x), _ptr(delta), _ptr(z), _ptr(A), _ptr(Bm), _ptr(Cm), _ptr(D), _ptr(dout),
                                   _ptr(chk), _ptr(dx), _ptr(dd), _ptr(dz), _ptr(dA), _ptr(dB), _ptr(dC), _ptr(dD),
                                   _ptr(ws), Bsz, L, ED, N, x.stride(1), delta.stride(1),
                                   z.stride(1) if z is not None else 0, dout.stride(1), lib.mmi_selscan_chunk(),
                                   _DT[dt], flags, _stream(x)), "mmi_selscan_bwd")
    launches += 2  # scan kernel + partial-reduction kernel
    return dx, dd, dz, dA, dB, dC, dD


class _SelectiveScan(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, delta, A, Bm, Cm, D, z, flags, grad_on):
        # grad_on = torch.is_grad_enabled() at the call site (inside forward() it is always off): under torch.no_grad() /
        # inference no checkpoints are written and nothing is saved, even though A_log / D are Parameters
        need = grad_on and any(ctx.needs_input_grad[:7])
        out, _, chk, saved = selscan_fwd_raw(x, delta, A, Bm, Cm, D, z=z, want_chk=need, flags=flags)
        ctx.flags = flags
        ctx.has_z = z is not None
        ctx.in_dtypes = tuple(None if t is None else t.dtype for t in (x, delta, A, Bm, Cm, D, z))
        if need:
            sx, sd, sz, sA, sB, sC, sD = saved
            ctx.save_for_backward(sx, sd, sA, sB, sC, sD, chk, *( [sz] if sz is not None else []))
        return out

    @staticmethod
    def backward(ctx, dout):
        sx, sd, sA, sB, sC, sD, chk, *rest = ctx.saved_tensors
        sz = rest[0] if rest else None
        dx, dd, dz, dA, dB, dC, dD = selscan_bwd_raw((sx, sd, sz, sA, sB, sC, sD), chk, dout, flags=ctx.flags)
        dts = ctx.in_dtypes
        cast = lambda g, i: None if g is None else g.to(dts[i])
        return cast(dx, 0), cast(dd, 1), cast(dA, 2), cast(dB, 3), cast(dC, 4), cast(dD, 5), (cast(dz, 6) if ctx.has_z else None), None, None


def selective_scan(x, delta, A, B, C, D, z=None, flags: int = 0, delta_softplus: bool = False):
    """Fused selective scan. Shapes as models/mamba.py:212-220: x, delta (B, L, ED); A (ED, N); B, C (B, L, N);
    D (ED).  Returns y (B, L, ED) = hs @ C + D * x, times silu(z) when `z` (B, L, ED) is given.
    delta_softplus=True: `delta` is the pre-activation dt_proj(.) and softplus (models/mamba.py:203) is applied inside
    the kernels (the returned gradient is then w.r.t. the pre-activation)."""
    if delta_softplus:
        flags |= _lib.FLAG_DELTA_SOFTPLUS
    return _SelectiveScan.apply(x, delta, A, B, C, D, z, flags, torch.is_grad_enabled())


class _CausalConv1dSiLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, silu):
        global launches
        lib = _lib.load()
        _require_cuda(x, "causal_conv1d")
        Bsz, L, ED = x.shape
        K = weight.shape[-1]
        if weight.numel() != ED * K or (bias is not None and tuple(bias.shape) != (ED,)):
            raise ValueError(f"causal_conv1d: weight {tuple(weight.shape)} / bias do not match a depthwise conv over {ED} channels")
        x = _rows(x)
        w = weight.detach().reshape(ED, K).float().contiguous()
        b = None if bias is None else bias.detach().float().contiguous()
        y = torch.empty((Bsz, L, ED), dtype=x.dtype, device=x.device)
        _lib.check(lib.mmi_causal_conv1d_fwd(_ptr(x), _ptr(w), _ptr(b), _ptr(y), Bsz, L, ED, K, x.stride(1), y.stride(1),
                                             _DT[x.dtype], int(silu), _stream(x)), "mmi_causal_conv1d_fwd")
        launches += 1
        ctx.save_for_backward(x, w, *([b] if b is not None else []))
        ctx.silu, ctx.wshape, ctx.wdtype, ctx.bdtype = silu, weight.shape, weight.dtype, None if bias is None else bias.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        global launches
        lib = _lib.load()
        x, w, *rest = ctx.saved_tensors
        b = rest[0] if rest else None
        Bsz, L, ED = x.shape
        K = w.shape[1]
        dy = _rows(dy.to(x.dtype))
        dx = torch.empty((Bsz, L, ED), dtype=x.dtype, device=x.device)
        dw = torch.empty_like(w)
        db = torch.empty(ED, dtype=torch.float32, device=x.device) if b is not None else None
        _lib.check(lib.mmi_causal_conv1d_bwd(_ptr(x), _ptr(w), _ptr(b), _ptr(dy), _ptr(dx), _ptr(dw), _ptr(db), Bsz, L, ED, K,
                                             x.stride(1), dy.stride(1), dx.stride(1), _DT[x.dtype], int(ctx.silu), _stream(x)),
                   "mmi_causal_conv1d_bwd")
        launches += 1
        return dx, dw.reshape(ctx.wshape).to(ctx.wdtype), (None if db is None else db.to(ctx.bdtype)), None


def causal_conv1d_silu(x, weight, bias=None, silu: bool = True):
    """Depthwise causal conv1d + SiLU on (B, L, ED) tokens: the x-branch prologue of MambaBlock.forward
    (models/mamba.py:176-180) without the transposes.  weight (ED, 1, K) as nn.Conv1d stores it, bias (ED) or None."""
    if x.dtype not in _DT:
        x = x.float()
    return _CausalConv1dSiLU.apply(x, weight, bias, silu)


class _RMSNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, eps, out_dtype):
        global launches
        lib = _lib.load()
        _require_cuda(x, "rmsnorm")
        C = x.shape[-1]
        x2 = x.reshape(-1, C)
        if x2.stride(1) != 1 or (x2.stride(0) * x2.element_size()) % 16 or x2.data_ptr() % 16:
            x2 = x2.contiguous()
        w = weight.detach().float().contiguous()
        y = torch.empty((x2.shape[0], C), dtype=out_dtype, device=x.device)
        _lib.check(lib.mmi_rmsnorm_fwd(_ptr(x2), _ptr(w), _ptr(y), x2.shape[0], C, x2.stride(0), y.stride(0), float(eps),
                                       _DT[x.dtype], _DT[out_dtype], _stream(x)), "mmi_rmsnorm_fwd")
        launches += 1
        ctx.save_for_backward(x2, w)
        ctx.eps, ctx.shape, ctx.wdtype, ctx.out_dtype = eps, x.shape, weight.dtype, out_dtype
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        global launches
        lib = _lib.load()
        x2, w = ctx.saved_tensors
        C = x2.shape[1]
        g = dy.to(ctx.out_dtype).reshape(-1, C)  # arrives in the output's dtype: read as is, no cast pass
        if g.stride(1) != 1 or (g.stride(0) * g.element_size()) % 16 or g.data_ptr() % 16:
            g = g.contiguous()
        dx = torch.empty((x2.shape[0], C), dtype=x2.dtype, device=x2.device)
        dw = torch.empty(C, dtype=torch.float32, device=x2.device)
        _lib.check(lib.mmi_rmsnorm_bwd(_ptr(x2), _ptr(w), _ptr(g), _ptr(dx), _ptr(dw), x2.shape[0], C, x2.stride(0), g.stride(0),
                                       dx.stride(0), float(ctx.eps), _DT[x2.dtype], _DT[ctx.out_dtype], _stream(x2)),
                   "mmi_rmsnorm_bwd")
        launches += 1
        return dx.view(ctx.shape), dw.to(ctx.wdtype), None, None


def rmsnorm(x, weight, eps: float = 1e-5):
    """RMSNorm.forward of models/mamba.py:356-366 as one kernel per direction (x: (..., C), weight: (C)).
    Under torch.autocast an fp32 input returns the autocast dtype (what the GEMM consuming it would cast to; same bits)."""
    if x.dtype not in _DT:
        x = x.float()
    out_dtype = x.dtype
    if x.dtype == torch.float32 and torch.is_autocast_enabled():
        # autocast runs norms in fp32 and casts their output for the 16-bit GEMM that follows: emit that dtype directly
        # (one rounding either way) and take the 16-bit gradient as it comes -- two cast passes over (B, L, C) saved
        ac = torch.get_autocast_dtype("cuda")
        if ac in (torch.bfloat16, torch.float16):
            out_dtype = ac
    return _RMSNorm.apply(x, weight, eps, out_dtype)


def _tok_gather(rgb, ir):
    global launches
    lib = _lib.load()
    B, C = rgb.shape[0], rgb.shape[1]
    HW = rgb[0, 0].numel()
    rgb, ir = rgb.contiguous(), ir.contiguous()
    tok = torch.empty((B, 2 * HW, C), dtype=rgb.dtype, device=rgb.device)
    _lib.check(lib.mmi_tokens_gather(_ptr(rgb), _ptr(ir), _ptr(tok), B, C, HW, _DT[rgb.dtype], _stream(rgb)), "mmi_tokens_gather")
    launches += 1
    return tok


def _tok_scatter(tok, shape):
    global launches
    lib = _lib.load()
    B, C = shape[0], shape[1]
    HW = tok.shape[1] // 2
    tok = tok.contiguous()
    rgb = torch.empty(shape, dtype=tok.dtype, device=tok.device)
    ir = torch.empty(shape, dtype=tok.dtype, device=tok.device)
    _lib.check(lib.mmi_tokens_scatter(_ptr(tok), _ptr(rgb), _ptr(ir), B, C, HW, _DT[tok.dtype], _stream(tok)), "mmi_tokens_scatter")
    launches += 1
    return rgb, ir


class _TokensGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb, ir):
        ctx.shape = rgb.shape
        return _tok_gather(rgb, ir)

    @staticmethod
    def backward(ctx, dtok):
        return _tok_scatter(dtok, ctx.shape)


class _TokensScatter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tok, shape):
        return _tok_scatter(tok, shape)

    @staticmethod
    def backward(ctx, drgb, dir_):
        return _tok_gather(drgb, dir_), None


def tokens_gather(rgb, ir):
    """Two NCHW maps (B, C, H, W) -> channels-last tokens (B, 2*H*W, C), VIS tokens first then IR (models/common.py:1338-1343)."""
    _require_cuda(rgb, "tokens_gather")
    if rgb.dtype not in _DT or rgb.shape != ir.shape or rgb.dtype != ir.dtype:
        raise RuntimeError("tokens_gather: rgb and ir must have the same shape and a float dtype")
    return _TokensGather.apply(rgb, ir)


def tokens_scatter(tok, shape):
    """Inverse of tokens_gather: tokens (B, 2*H*W, C) -> (rgb, ir) of `shape` = (B, C, H, W) (models/common.py:1352-1366)."""
    _require_cuda(tok, "tokens_scatter")
    return _TokensScatter.apply(tok, tuple(shape))


def _resample(fn, src, out_shape, src_is_big):
    """the four entry points share one signature: (src, dst, B*C, H, W, hs, ws) with (H, W) the map, (hs, ws) the anchors"""
    global launches
    lib = _lib.load()
    src = src.contiguous()
    dst = torch.empty(out_shape, dtype=src.dtype, device=src.device)
    big, small = (src, dst) if src_is_big else (dst, src)
    _lib.check(getattr(lib, fn)(_ptr(src), _ptr(dst), src.shape[0] * src.shape[1], big.shape[2], big.shape[3], small.shape[2],
                                small.shape[3], _DT[src.dtype], _stream(src)), fn)
    launches += 1
    return dst


class _AvgPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, size):
        ctx.shape = x.shape
        return _resample("mmi_avgpool_fwd", x, (*x.shape[:2], *size), True)

    @staticmethod
    def backward(ctx, dy):
        return _resample("mmi_avgpool_bwd", dy, ctx.shape, False), None


class _UpsampleBilinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, size):
        ctx.shape = x.shape
        return _resample("mmi_upsample_bilinear_fwd", x, (*x.shape[:2], *size), False)

    @staticmethod
    def backward(ctx, dy):
        return _resample("mmi_upsample_bilinear_bwd", dy, ctx.shape, True), None


def adaptive_avg_pool(x, size):
    """nn.AdaptiveAvgPool2d(size) of models/common.py:324-325 on (B, C, H, W); one pass over the map per direction."""
    _require_cuda(x, "adaptive_avg_pool")
    if x.dtype not in _DT:
        x = x.float()
    return _AvgPool.apply(x, tuple(int(v) for v in size))


def upsample_bilinear(x, size):
    """F.interpolate(x, size=size, mode='bilinear') (align_corners=False) of models/common.py:540-543."""
    _require_cuda(x, "upsample_bilinear")
    if x.dtype not in _DT:
        x = x.float()
    return _UpsampleBilinear.apply(x, tuple(int(v) for v in size))
