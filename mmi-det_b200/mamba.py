"""Host-side mirror of models/mamba.py for the B200 path, plus the VIS+IR fusion module and the plug-in helper.

Same class names, constructor signatures, parameter names and shapes as the reference, so a reference state_dict loads
unchanged and pickled / deep-copied modules behave the same (no ctypes handles or streams are stored on a
module; the library is loaded lazily by the operator layer):

    MambaConfig      models/mamba.py:30-54     RMSNorm        models/mamba.py:356-366
    MambaBlock       models/mamba.py:117-353   ResidualBlock  models/mamba.py:89-115      Mamba   models/mamba.py:56-87

What differs is only where the arithmetic runs: `MambaBlock.selective_scan` and the SiLU gate (models/mamba.py:212-233,
184-186) go through the fused sm_100a kernels (ops.selective_scan); the four projections stay nn.Linear (cuBLAS); the
depthwise causal conv + SiLU prologue (models/mamba.py:176-180) runs as one channels-last kernel.  Unlike the reference the block is
half/bf16-clean: inputs of any float dtype are accepted and the input dtype is returned.

`MambaFusion` is the cross-modal block the north_star names (the reference ships none): it keeps the I/O
contract of the reference's `GPT` fusion transformer (models/common.py:1270-1370: ctor called with d_model only,
forward([rgb, ir]) -> (rgb_out, ir_out) with input shapes preserved, VIS tokens first then IR tokens) so that
`install(fusion=True)` can bind it to the YAML name `GPT` that the unchanged parse_model looks up (models/yolo_test.py:560,
600-602).  CUDA only: there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


@dataclass
class MambaConfig:
    d_model: int
    n_layers: int
    dt_rank: Union[int, str] = "auto"
    d_state: int = 16
    expand_factor: int = 2
    d_conv: int = 4
    dt_min: float = 0.001
    dt_max: float = 0.1
    dt_init: str = "random"
    dt_scale: float = 1.0
    dt_init_floor = 1e-4
    bias: bool = False
    conv_bias: bool = True
    pscan: bool = True  # kept for signature compatibility; both settings run the same fused kernel

    def __post_init__(self):
        self.d_inner = self.expand_factor * self.d_model
        if self.dt_rank == "auto":
            self.dt_rank = math.ceil(self.d_model / 16)


class RMSNorm(nn.Module):
    def __init__(self, d_model: int, eps: float = 1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(d_model))

    def forward(self, x):
        # one kernel per direction instead of five elementwise passes; widths up to 4096 in multiples of 8 (every d_model
        # of the detector family).  No eager fallback: other widths / CPU tensors raise in the operator.
        return ops.rmsnorm(x, self.weight, self.eps)


class MambaBlock(nn.Module):
    def __init__(self, config: MambaConfig):
        super().__init__()
        self.config = config
        c = config
        self.in_proj = nn.Linear(c.d_model, 2 * c.d_inner, bias=c.bias)
        self.conv1d = nn.Conv1d(c.d_inner, c.d_inner, kernel_size=c.d_conv, bias=c.conv_bias, groups=c.d_inner,
                                padding=c.d_conv - 1)
        self.x_proj = nn.Linear(c.d_inner, c.dt_rank + 2 * c.d_state, bias=False)
        self.dt_proj = nn.Linear(c.dt_rank, c.d_inner, bias=True)
        # same initial distribution as the reference (models/mamba.py:137-160): dt weights ~ U(+-dt_rank^-0.5 * scale) or
        # constant; dt bias = softplus^-1 of a log-uniform step in [dt_min, dt_max]; A = -(1..N) per channel; D = 1
        std = c.dt_rank ** -0.5 * c.dt_scale
        if c.dt_init == "constant":
            nn.init.constant_(self.dt_proj.weight, std)
        elif c.dt_init == "random":
            nn.init.uniform_(self.dt_proj.weight, -std, std)
        else:
            raise NotImplementedError(c.dt_init)
        lo, hi = math.log(c.dt_min), math.log(c.dt_max)
        dt = torch.exp(torch.rand(c.d_inner) * (hi - lo) + lo).clamp(min=c.dt_init_floor)
        with torch.no_grad():
            self.dt_proj.bias.copy_(dt + torch.log(-torch.expm1(-dt)))
        self.A_log = nn.Parameter(torch.log(torch.arange(1, c.d_state + 1, dtype=torch.float32).repeat(c.d_inner, 1)))
        self.D = nn.Parameter(torch.ones(c.d_inner))
        self.out_proj = nn.Linear(c.d_inner, c.d_model, bias=c.bias)

    def _apply(self, fn, recurse=True):
        """`.half()` / `.bfloat16()` (detect_twostream.py:45, test.py:68) keep A_log and D in fp32: both are only ever used
        up-cast (`A_log.float()`, `D.float()`, models/mamba.py:196-197), and rounding log(1..N) to 16 bits would move every row
        of A off the S4D-real geometric form -- the model would drift from its fp32 self and the scan would need N exponentials
        per step instead of one.  Device moves and fp32 / fp64 casts apply to them as to any parameter."""
        keep = {n: getattr(self, n).data for n in ("A_log", "D") if getattr(self, n, None) is not None}
        out = super()._apply(fn, recurse)
        for n, v in keep.items():
            p = getattr(self, n)
            if p.dtype in (torch.float16, torch.bfloat16) and v.dtype == torch.float32:
                p.data = v.to(p.device)
        return out

    # -- the hot path -----------------------------------------------------------------------------------------
    def selective_scan(self, x, delta, A, B, C, D, z=None, delta_softplus=False):
        """models/mamba.py:212-233 (and :235-265, same maths) on the fused kernel; `z` additionally fuses the gate,
        `delta_softplus` the softplus of models/mamba.py:203."""
        return ops.selective_scan(x, delta, A, B, C, D, z=z, delta_softplus=delta_softplus)

    selective_scan_seq = selective_scan

    def ssm(self, x, z=None):
        c = self.config
        A = -torch.exp(self.A_log.float())
        dbc = self.x_proj(x)
        delta, B, C = torch.split(dbc, [c.dt_rank, c.d_state, c.d_state], dim=-1)
        # dt_proj on the low-rank slice as ONE addmm (bias in the GEMM epilogue): F.linear on this strided view otherwise
        # runs matmul + a separate broadcast bias pass over (B, L, ED).  softplus(dt_proj(.)) of models/mamba.py:203 is
        # applied inside the scan kernels (no elementwise pass, no cast).
        dtp = self.dt_proj
        if dtp.bias is not None and delta.dim() == 3:
            delta_pre = torch.addmm(dtp.bias, delta.reshape(-1, c.dt_rank), dtp.weight.t()).view(*delta.shape[:-1], -1)
        else:
            delta_pre = dtp(delta)
        return self.selective_scan(x, delta_pre, A, B, C, self.D.float(), z=z, delta_softplus=True)

    def forward(self, x, residual=None):
        """models/mamba.py:165-189.  `residual` (same shape and dtype as the output) is added in out_proj's GEMM epilogue
        (addmm, beta = 1): ResidualBlock's `mixer(norm(x)) + x` (models/mamba.py:101) without a separate pass over (B, L, D)."""
        xz = self.in_proj(x)
        xs, z = xz.chunk(2, dim=-1)  # views of one GEMM output: passed to the kernel by row pitch, never copied
        if self.config.d_conv > 4 or self.conv1d.groups != self.conv1d.in_channels:
            raise RuntimeError(f"mmidet_b200.MambaBlock: d_conv={self.config.d_conv} is outside the fused causal conv kernel "
                               "(depthwise, kernel size <= 4; the reference default is 4) -- there is no eager fallback")
        # depthwise causal conv + SiLU on the channels-last tokens (no transposes, models/mamba.py:176-180)
        xs = ops.causal_conv1d_silu(xs, self.conv1d.weight, self.conv1d.bias)
        y = self.ssm(xs, z=z)  # = ssm(x) * silu(z): the gate of models/mamba.py:184-186 is fused into the scan
        y = y.to(xz.dtype)
        op = self.out_proj
        if residual is not None and op.bias is None and residual.dtype == y.dtype and not torch.is_autocast_enabled():
            return torch.addmm(residual.reshape(-1, residual.shape[-1]), y.reshape(-1, y.shape[-1]), op.weight.t()).view(residual.shape)
        out = op(y)
        return out if residual is None else out + residual

    # -- single-token recurrent inference (models/mamba.py:289-353): same cache contract (h (B, ED, N) or None, inputs
    #    (B, ED, d_conv-1)); the conv window and the one-step scan run on the same kernels as forward() (L = d_conv / L = 1,
    #    state passed through h0 -> hT), the projections on cuBLAS
    def step(self, x, cache):
        h, inputs = cache
        xz = self.in_proj(x)
        xs, z = xz.chunk(2, dim=1)
        xc = xs.unsqueeze(2)
        win = torch.cat([inputs.to(xs.dtype), xc], dim=2).transpose(1, 2).contiguous()  # (B, d_conv, ED): the causal window
        xs = ops.causal_conv1d_silu(win, self.conv1d.weight, self.conv1d.bias)[:, -1]
        y, h = self.ssm_step(xs, h, z=z)
        out = self.out_proj(y.to(xz.dtype))
        return out, (h, torch.cat([inputs[:, :, 1:], xc.to(inputs.dtype)], dim=2))

    def ssm_step(self, x, h, z=None):
        """models/mamba.py:322-353: one step of the recurrence from state `h` (None = zeros); returns (y, h_new).
        `z` (B, ED) additionally fuses the gate y * silu(z) of step() (models/mamba.py:311-313)."""
        c = self.config
        A = -torch.exp(self.A_log.float())
        delta, B, C = torch.split(self.x_proj(x), [c.dt_rank, c.d_state, c.d_state], dim=-1)
        delta_pre = self.dt_proj(delta)
        one = lambda t: t.unsqueeze(1).contiguous()
        y, hT, _, _ = ops.selscan_fwd_raw(one(x), one(delta_pre), A, one(B), one(C), self.D.float(),
                                          z=None if z is None else one(z), h0=h, want_state=True,
                                          flags=ops._lib.FLAG_DELTA_SOFTPLUS)
        return y[:, 0], hT


class ResidualBlock(nn.Module):
    def __init__(self, config: MambaConfig):
        super().__init__()
        self.mixer = MambaBlock(config)
        self.norm = RMSNorm(config.d_model)

    def forward(self, x):
        return self.mixer(self.norm(x), residual=x)  # + x folded into out_proj's epilogue where dtypes allow

    def step(self, x, cache):
        out, cache = self.mixer.step(self.norm(x), cache)
        return out + x, cache


class Mamba(nn.Module):
    def __init__(self, config: MambaConfig):
        super().__init__()
        self.config = config
        self.layers = nn.ModuleList([ResidualBlock(config) for _ in range(config.n_layers)])

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return x

    def step(self, x, caches):
        for i, layer in enumerate(self.layers):
            x, caches[i] = layer.step(x, caches[i])
        return x, caches


def _is_channels_last(t):
    return t.dim() == 4 and t.stride(1) == 1 and t.stride(3) == t.shape[1] and t.stride(2) == t.shape[3] * t.shape[1] \
        and t.stride(0) == t.shape[2] * t.shape[3] * t.shape[1]


class MambaFusion(nn.Module):
    """Cross-modal fusion with the `GPT` contract (models/common.py:1270-1370).

    forward([rgb, ir]) with rgb, ir (B, C, H, W): both maps are flattened to tokens at FULL resolution (no 8x8 pooling:
    L = H*W per modality, i.e. L = 6400 at P3 / 640 px, the BASELINE shape), concatenated along the sequence axis as
    VIS tokens then IR tokens (models/common.py:1340-1343), run through `n_layer` ResidualBlocks (RMSNorm -> MambaBlock ->
    + x, models/mamba.py:89-102) so the state carried out of the VIS half conditions the IR half, and split back.
    `block_cls` lets a test build the very same module on the reference's pure-PyTorch ResidualBlock (the logits oracle).
    Extra ctor arguments of GPT (h, block_exp, vert_anchors, ... ) are accepted and ignored."""

    def __init__(self, d_model, h=8, block_exp=4, n_layer=1, vert_anchors=8, horz_anchors=8, embd_pdrop=0.1, attn_pdrop=0.1,
                 resid_pdrop=0.1, block_cls=None, config_cls=None):
        super().__init__()
        cfg = (config_cls or MambaConfig)(d_model=d_model, n_layers=n_layer)
        self.n_embd = d_model
        self.reference_glue = block_cls is not None
        self.layers = nn.ModuleList([(block_cls or ResidualBlock)(cfg) for _ in range(n_layer)])

    def forward(self, x):
        rgb, ir = x[0], x[1]
        B, C, H, W = rgb.shape
        if self.reference_glue:
            # the logits oracle: the SAME module built on the reference's pure-PyTorch blocks (block_cls=) also gets the
            # reference-style torch glue, so that arm contains nothing of ours (tests / the PyTorch-GPU comparison arm)
            tok = torch.cat([rgb.flatten(2), ir.flatten(2)], dim=2).transpose(1, 2).contiguous()
            for layer in self.layers:
                tok = layer(tok)
            out = tok.transpose(1, 2).reshape(B, C, 2, H, W)
            return out[:, :, 0].contiguous(), out[:, :, 1].contiguous()
        if ir.dtype != rgb.dtype:
            ir = ir.to(rgb.dtype)
        if rgb.is_cuda and C > 1 and _is_channels_last(rgb) and _is_channels_last(ir):
            # a channels_last backbone hands over maps that already ARE token rows (B, HW, C): the gather is the concatenation
            # of two contiguous blocks and the way back is two views of the token tensor -- no transpose in either direction
            tok = torch.cat([rgb.permute(0, 2, 3, 1).reshape(B, H * W, C), ir.permute(0, 2, 3, 1).reshape(B, H * W, C)], dim=1)
            for layer in self.layers:
                tok = layer(tok)
            out = tok.view(B, 2, H, W, C)
            return out[:, 0].permute(0, 3, 1, 2), out[:, 1].permute(0, 3, 1, 2)
        tok = ops.tokens_gather(rgb, ir)  # (B, 2HW, C), VIS tokens then IR tokens: one tiled transpose (raises off-GPU)
        for layer in self.layers:
            tok = layer(tok)
        return ops.tokens_scatter(tok, (B, C, H, W))


def install(ref_models=None, scan=True, pscan=True, ffm=True, fusion=False):
    """Bind the B200 path into an imported, UNMODIFIED reference tree -- the stubs INTEGRATION.md shows:

        import models.mamba, models.common, models.yolo_test      # the reference, untouched
        import mmidet_b200.mamba as M; M.install(fusion=True)

    scan    MambaBlock.selective_scan / selective_scan_seq  -> fused kernel (class attribute patch; ctor untouched)
    pscan   models.mamba.pscan (= PScan.apply)               -> mmidet_b200.pscan.pscan
    ffm     models.common.extract_frequency2, Seperation_loss -> mmidet_b200.ffm; GPT1_fourier.forward -> ffm.fourier_forward,
            GPT1.forward -> ffm.gpt1_forward
            (class attribute patch: the module keeps its ctor, parameters and state_dict)
    fusion  models.yolo_test.GPT                              -> MambaFusion (YAML rows naming GPT then build it)
    Returns the list of replaced attributes so a caller can restore them."""
    import importlib

    from . import ffm as _ffm
    from . import pscan as _pscan
    saved = []

    def mod(name):
        return importlib.import_module(name) if ref_models is None else getattr(ref_models, name.split(".")[-1])

    def swap(obj, attr, new):
        saved.append((obj, attr, getattr(obj, attr)))
        setattr(obj, attr, new)

    if scan or pscan:
        mm = mod("models.mamba")
        if scan:
            fused = lambda self, x, delta, A, B, C, D: ops.selective_scan(x, delta, A, B, C, D)  # noqa: E731
            swap(mm.MambaBlock, "selective_scan", fused)
            swap(mm.MambaBlock, "selective_scan_seq", fused)
        if pscan:
            swap(mm, "pscan", _pscan.pscan)
    if ffm:
        mc = mod("models.common")
        swap(mc, "extract_frequency2", _ffm.extract_frequency2)
        swap(mc, "Seperation_loss", _ffm.separation_loss)
        if hasattr(mc, "GPT1_fourier"):
            swap(mc.GPT1_fourier, "forward", _ffm.fourier_forward)
        if hasattr(mc, "GPT1"):
            swap(mc.GPT1, "forward", _ffm.gpt1_forward)
    if fusion:
        swap(mod("models.yolo_test"), "GPT", MambaFusion)
    return saved


def uninstall(saved):
    for obj, attr, old in reversed(saved):
        setattr(obj, attr, old)
