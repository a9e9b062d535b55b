#!/usr/bin/env python
"""tests/golden/make_golden.py -- regenerate the golden fixtures from the UNMODIFIED reference.

Run in the build container only (needs /root/reference; the GPU box has no reference tree):

    python tests/golden/make_golden.py

Imports models/pscan.py, models/mamba.py and models/common.py from the reference (CPU, fp32 / fp64),
feeds seeded synthetic inputs and stores inputs + outputs as small .npz files next to this script.
Nothing here is copied from the reference; it is executed as a black box.
"""
import os
import sys
import unittest.mock as mock

import numpy as np
import torch

REF = os.environ.get("MMIDET_REF", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

for n in ("matplotlib", "matplotlib.pyplot", "seaborn", "thop", "torchsummary"):  # absent in this image
    sys.modules.setdefault(n, mock.MagicMock())
sys.path.insert(0, REF)

from models.pscan import pscan  # noqa: E402
from models.mamba import MambaBlock, MambaConfig, ResidualBlock  # noqa: E402
import models.common as C  # noqa: E402


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = v
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def scan_inputs(B, L, ED, N, seed, dtype=torch.float32, random_A=False):
    """SURVEY 8(d) config-1 scan-only inputs."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, L, ED, generator=g, dtype=dtype)
    delta = torch.nn.functional.softplus(torch.randn(B, L, ED, generator=g, dtype=dtype) - 3.0)
    z = torch.randn(B, L, ED, generator=g, dtype=dtype)
    Bm = torch.randn(B, L, N, generator=g, dtype=dtype)
    Cm = torch.randn(B, L, N, generator=g, dtype=dtype)
    if random_A:
        A_log = torch.randn(ED, N, generator=g, dtype=dtype) * 0.7 + 0.5
        D = torch.randn(ED, generator=g, dtype=dtype)
    else:  # mamba.py:158-160 default init
        A_log = torch.log(torch.arange(1, N + 1, dtype=torch.float32).repeat(ED, 1)).to(dtype)
        D = torch.ones(ED, dtype=dtype)
    A = -torch.exp(A_log.float()).to(dtype)
    return x, delta, z, A, Bm, Cm, D


def gen_pscan():
    """models/pscan.py fwd + bwd on pow2 and non-pow2 L (fp32), plus an fp64 run."""
    for tag, (B, L, D, N), dt in (("pow2", (2, 64, 8, 16), torch.float32), ("ragged", (2, 37, 8, 16), torch.float32),
                                  ("tiny", (1, 3, 4, 16), torch.float32), ("f64", (1, 50, 4, 16), torch.float64)):
        g = torch.Generator().manual_seed(11)
        A = torch.rand(B, L, D, N, generator=g, dtype=dt) * 0.9 + 0.05
        X = torch.randn(B, L, D, N, generator=g, dtype=dt)
        gH = torch.randn(B, L, D, N, generator=g, dtype=dt)
        A.requires_grad_(True)
        X.requires_grad_(True)
        H = pscan(A, X)
        gA, gX = torch.autograd.grad(H, (A, X), gH)
        save(f"pscan_{tag}", A=A, X=X, gH=gH, H=H, gA=gA, gX=gX)


def gen_selscan():
    """MambaBlock.selective_scan (pscan path), selective_scan_seq, gate, and autograd gradients."""
    for tag, (B, L, ED, N), rnd in (("init", (2, 96, 24, 16), False), ("randA", (2, 75, 16, 16), True),
                                    ("short", (1, 5, 8, 16), True)):
        x, delta, z, A, Bm, Cm, D = scan_inputs(B, L, ED, N, seed=3, random_A=rnd)
        blk = MambaBlock(MambaConfig(d_model=ED // 2, n_layers=1, d_state=N))
        leaves = [t.clone().requires_grad_(True) for t in (x, delta, z, A, Bm, Cm, D)]
        xr, dr, zr, Ar, Br, Cr, Dr = leaves
        y_pscan = blk.selective_scan(xr, dr, Ar, Br, Cr, Dr)            # mamba.py:212
        y_seq = blk.selective_scan_seq(xr, dr, Ar, Br, Cr, Dr)          # mamba.py:235
        out = y_pscan * torch.nn.functional.silu(zr)                    # mamba.py:184-186
        g = torch.Generator().manual_seed(5)
        dout = torch.randn(B, L, ED, generator=g)
        grads = torch.autograd.grad(out, leaves, dout)
        save(f"selscan_{tag}", x=x, delta=delta, z=z, A=A, Bm=Bm, Cm=Cm, D=D, y_pscan=y_pscan, y_seq=y_seq, out=out,
             dout=dout, dx=grads[0], ddelta=grads[1], dz=grads[2], dA=grads[3], dB=grads[4], dC=grads[5], dD=grads[6])


def gen_block():
    """Whole MambaBlock / ResidualBlock forward + input gradient with the reference's own init."""
    torch.manual_seed(0)
    cfg = MambaConfig(d_model=16, n_layers=1)
    blk = ResidualBlock(cfg)
    x = torch.randn(2, 48, 16, requires_grad=True)
    y_mixer = blk.mixer(x)
    y_res = blk(x)
    g = torch.randn(2, 48, 16)
    (gx,) = torch.autograd.grad(y_res, x, g, retain_graph=True)
    pgrads = torch.autograd.grad(y_res, list(blk.parameters()), g)
    arrs = {"sd." + k: v for k, v in blk.state_dict().items()}
    arrs.update({"pg." + n: pg for (n, _), pg in zip(blk.named_parameters(), pgrads)})
    save("mamba_block", x=x, y_mixer=y_mixer, y_res=y_res, g=g, gx=gx, **arrs)


def gen_ffm():
    """extract_frequency2 at the sizes SURVEY F3 probed (negative-slice quirk) + fourier_transform + Seperation_loss."""
    for hw in (8, 16, 20, 7, (8, 12)):
        h, w = (hw, hw) if isinstance(hw, int) else hw
        g = torch.Generator().manual_seed(h * 100 + w)
        img = torch.randn(2, 3, h, w, generator=g)
        low, high = C.extract_frequency2(img)
        fs = C.fourier_transform(img)
        save(f"ffm_{h}x{w}", img=img, low=low, high=high, fs_re=fs.real, fs_im=fs.imag)
    g = torch.Generator().manual_seed(9)
    M = torch.rand(36, 64, generator=g)
    save("seploss", M=M, loss=C.Seperation_loss(M))


if __name__ == "__main__":
    torch.set_num_threads(4)
    gen_pscan()
    gen_selscan()
    gen_block()
    gen_ffm()
