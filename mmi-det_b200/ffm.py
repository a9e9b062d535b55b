"""Fusion Focus Module Fourier step -- drop-ins for the helpers of models/common.py.

extract_frequency2(image) -> (low, high): same contract as models/common.py:37-69 (both outputs torch.float16,
real, shape of `image`), including the negative-slice wrap of the reference's masks (:44-56).
separation_loss(M): models/common.py:128-139 (`Seperation_loss`) in closed form.
pattern_tokens(...): the pattern path of GPT1_fourier.forward between pooling and transformer (models/common.py:440-516)
as one launch per direction (csrc/ffm_pattern.cu); fourier_forward is the forward of GPT1_fourier (:357-552) rebuilt on it,
bound onto the reference class by mamba.install(ffm=True).
Everything runs on csrc/ffm*.cu through the C ABI; CUDA tensors only."""
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
    if image.requires_grad and torch.is_grad_enabled():
        # the reference's torch.fft version is differentiable; this kernel is a value-only drop-in (the detector detaches
        # everything downstream of it, models/yolo_test.py:226-230): refuse rather than return a silent zero gradient
        raise RuntimeError("mmidet_b200.extract_frequency2 is not differentiable: call it under torch.no_grad() or on a "
                           "detached tensor (GPT1_fourier.forward goes through ffm.pattern_tokens, which is)")
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


def extract_frequency(image: torch.Tensor, threshold: int = 30):
    """models/common.py:72-93 (no caller in the reference): zero the central 2*threshold block of the shifted spectrum
    ("low"), the rest is "high"; both returned as the REAL part in fp16, which is what `.half()` of a complex tensor yields.
    The FFT is the library call (cuFFT through torch.fft), the split is two masked writes."""
    if not image.is_cuda:
        raise RuntimeError("mmidet_b200.extract_frequency: CUDA tensor required (no CPU path)")
    fs = fourier_transform(image.float()).real
    H, W = image.shape[-2:]
    ch, cw = H // 2, W // 2
    low = fs.clone()
    low[:, :, ch - threshold:ch + threshold, cw - threshold:cw + threshold] = 0  # Python slice semantics, as the reference
    return low.half(), (fs - low).half()


def separation_loss(M: torch.Tensor) -> torch.Tensor:
    """models/common.py:128-139 for M (l, K): sum_{i<j} M_i . M_j / (l (l - 1)), as a 0-d fp32 tensor."""
    if not M.is_cuda:
        raise RuntimeError("mmidet_b200.separation_loss: CUDA tensor required (no CPU path)")
    if M.requires_grad and torch.is_grad_enabled():
        raise RuntimeError("mmidet_b200.separation_loss returns a value only (the detector detaches it, "
                           "models/yolo_test.py:230): call it under torch.no_grad() or on a detached tensor")
    lib = _lib.load()
    Mc = M.detach().float().contiguous()
    out = torch.empty(1, dtype=torch.float32, device=M.device)
    _lib.check(lib.mmi_separation_loss(_ops._ptr(Mc), _ops._ptr(out), Mc.shape[0], Mc.shape[1], _ops._stream(Mc)),
               "mmi_separation_loss")
    _ops.launches += 1
    return out[0]


Seperation_loss = separation_loss  # the reference's spelling (models/common.py:128)


def _pattern_fwd(vis, ir, w1, w2, high=True):
    lib = _lib.load()
    B, C, H, W = vis.shape
    tok = torch.empty((B, 2 * H * W, C), dtype=vis.dtype, device=vis.device)
    rows = torch.empty((18 * B, H * W), dtype=torch.float32, device=vis.device)
    loss = torch.empty(1, dtype=torch.float32, device=vis.device)
    ws = torch.empty(lib.mmi_ffm_pattern_ws_bytes(B, C, H * W), dtype=torch.uint8, device=vis.device) if high else None
    _lib.check(lib.mmi_ffm_pattern_fwd(_ops._ptr(vis), _ops._ptr(ir), _ops._ptr(w1), _ops._ptr(w2), _ops._ptr(tok),
                                       _ops._ptr(rows), _ops._ptr(loss), _ops._ptr(ws), B, C, H, W, _ops._DT[vis.dtype],
                                       _ops._stream(vis)), "mmi_ffm_pattern_fwd")
    _ops.launches += 3 if high else 2
    return tok, rows, loss[0]


class _PatternTokens(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vis, ir, conv1_w, conv2_w, high):
        C = vis.shape[1]
        w1 = conv1_w.detach().reshape(8, C).float().contiguous()
        w2 = conv2_w.detach().reshape(C, 8).float().contiguous()
        vis, ir = vis.contiguous(), ir.contiguous()
        tok, rows, loss = _pattern_fwd(vis, ir, w1, w2, high)
        ctx.save_for_backward(vis, ir, w1, w2, rows)
        ctx.wshape = (conv1_w.shape, conv2_w.shape, conv1_w.dtype, conv2_w.dtype)
        ctx.mark_non_differentiable(loss)
        return tok, loss

    @staticmethod
    def backward(ctx, dtok, _dloss):
        vis, ir, w1, w2, rows = ctx.saved_tensors
        lib = _lib.load()
        B, C, P = vis.shape[0], vis.shape[1], vis.shape[2] * vis.shape[3]
        dtok = dtok.to(vis.dtype).contiguous()
        dvis, dir_ = torch.empty_like(vis), torch.empty_like(ir)
        dw1 = torch.empty((8, C), dtype=torch.float32, device=vis.device)
        dw2 = torch.empty((C, 8), dtype=torch.float32, device=vis.device)
        ws = torch.empty(lib.mmi_ffm_pattern_ws_bytes(B, C, P), dtype=torch.uint8, device=vis.device)
        _lib.check(lib.mmi_ffm_pattern_bwd(_ops._ptr(vis), _ops._ptr(ir), _ops._ptr(dtok), _ops._ptr(rows), _ops._ptr(w1),
                                           _ops._ptr(w2), _ops._ptr(dvis), _ops._ptr(dir_), _ops._ptr(dw1), _ops._ptr(dw2),
                                           _ops._ptr(ws), B, C, P, _ops._DT[vis.dtype], _ops._stream(vis)),
                   "mmi_ffm_pattern_bwd")
        _ops.launches += 2
        s1, s2, d1, d2 = ctx.wshape
        return dvis, dir_, dw1.reshape(s1).to(d1), dw2.reshape(s2).to(d2), None


def pattern_tokens(vis: torch.Tensor, ir: torch.Tensor, conv1_weight: torch.Tensor, conv2_weight: torch.Tensor, high: bool = True):
    """Pattern path of GPT1_fourier.forward (models/common.py:440-516) on the pooled maps vis, ir (B, C, h, w):
    -> (token_embeddings (B, 2hw, C), pattenLoss 0-d fp32).  conv1_weight (8, C, 1, 1), conv2_weight (C, 8, 1, 1).
    high=False: the same path of GPT1.forward (models/common.py:218-262), which has no Fourier branch (loss over 16B rows).
    Gradients flow through the tokens to vis, ir and both weights; the loss is a value only (the reference detaches
    it, models/yolo_test.py:230)."""
    if not (vis.is_cuda and ir.is_cuda):
        raise RuntimeError("mmidet_b200.pattern_tokens: CUDA tensors required (no CPU path)")
    if vis.shape != ir.shape or vis.dim() != 4:
        raise ValueError("pattern_tokens: vis and ir must be (B, C, h, w) maps of one shape")
    if conv1_weight.shape[0] != 8 or conv1_weight.numel() != 8 * vis.shape[1] or conv2_weight.numel() != 8 * vis.shape[1]:
        raise ValueError("pattern_tokens: conv1 must map C -> 8 and conv2 8 -> C (models/common.py:330-336)")
    if vis.dtype not in _ops._DT:
        vis = vis.float()
    return _PatternTokens.apply(vis, ir.to(vis.dtype), conv1_weight, conv2_weight, bool(high))


def _focus_forward(self, x, high):
    rgb_fea, ir_fea = x[0], x[1]
    assert rgb_fea.shape[0] == ir_fea.shape[0]
    bs, c, h, w = rgb_fea.shape
    anchors = (self.vert_anchors, self.horz_anchors)
    tok, self.pattenLoss = pattern_tokens(_ops.adaptive_avg_pool(rgb_fea, anchors), _ops.adaptive_avg_pool(ir_fea, anchors),
                                          self.conv1.weight, self.conv2.weight, high=high)
    t = self.drop(self.pos_emb + tok)
    t = self.ln_f(self.trans_blocks(t))
    rgb_out, ir_out = _ops.tokens_scatter(t, (bs, self.n_embd, self.vert_anchors, self.horz_anchors))
    return _ops.upsample_bilinear(rgb_out, (h, w)), _ops.upsample_bilinear(ir_out, (h, w)), self.pattenLoss


def fourier_forward(self, x):
    """Replacement for GPT1_fourier.forward (models/common.py:357-552) running on `self`'s own sub-modules and
    parameters (conv1, conv2, pos_emb, drop, trans_blocks, ln_f): the transformer and LayerNorm stay library calls;
    the adaptive pooling to the anchor grid and the bilinear upsample back are the resampling kernels
    (ops.adaptive_avg_pool / upsample_bilinear); the Fourier split, both conv1+sigmoid branches, the separation loss,
    conv2 * fea and the token layout are pattern_tokens; tokens go back to two maps with the token-scatter kernel.
    Returns (rgb_fea_out, ir_fea_out, pattenLoss) and sets self.pattenLoss like the reference."""
    return _focus_forward(self, x, True)


def gpt1_forward(self, x):
    """Replacement for GPT1.forward (models/common.py:196-298), the sibling of GPT1_fourier without the Fourier branch:
    same pooling / pattern path / transformer / upsample, separation loss over [M_vis; M_ir] only (:218-239)."""
    return _focus_forward(self, x, False)
