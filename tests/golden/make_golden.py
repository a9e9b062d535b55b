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


def gen_step():
    """MambaBlock.step / ssm_step (models/mamba.py:289-353) through ResidualBlock.step over 8 tokens from the empty cache,
    next to forward() on the same 8-token prefix (the two must agree: recurrent == parallel form)."""
    torch.manual_seed(4)
    cfg = MambaConfig(d_model=16, n_layers=1)
    blk = ResidualBlock(cfg).eval()
    with torch.no_grad():
        blk.mixer.A_log.add_(torch.randn_like(blk.mixer.A_log) * 0.3)  # leave the S4D-real init: general-A path too
        blk.mixer.D.normal_(1.0, 0.2)
    B, T = 3, 8
    x = torch.randn(B, T, 16)
    cache = (None, torch.zeros(B, cfg.d_inner, cfg.d_conv - 1))
    ys, hs = [], []
    with torch.no_grad():
        for t in range(T):
            y, cache = blk.step(x[:, t], cache)
            ys.append(y)
            hs.append(cache[0])
        y_fwd = blk(x)
    arrs = {"sd." + k: v for k, v in blk.state_dict().items()}
    save("mamba_step", x=x, y_step=torch.stack(ys, 1), h_step=torch.stack(hs, 1), inputs_last=cache[1], y_fwd=y_fwd, **arrs)


def gen_ffm():
    """extract_frequency2 at the sizes SURVEY F3 probed (negative-slice quirk) + fourier_transform + Seperation_loss."""
    for hw in (8, 16, 20, 7, (8, 12), 80, 160, (96, 72)):
        h, w = (hw, hw) if isinstance(hw, int) else hw
        g = torch.Generator().manual_seed(h * 100 + w)
        big = max(h, w) > 64
        img = torch.randn(1 if big else 2, 2 if big else 3, h, w, generator=g)
        low, high = C.extract_frequency2(img)
        fs = C.fourier_transform(img)
        extra = {}
        if h == w and h in (8, 80):  # extract_frequency (common.py:72-93, no caller in the reference): fixed threshold 30
            lo1, hi1 = C.extract_frequency(img)
            extra = dict(ef_low=lo1.float(), ef_high=hi1.float())
        if big:  # keep the large fixtures small: spectrum omitted (fourier_transform is pinned at the small sizes)
            save(f"ffm_{h}x{w}", img=img.half(), low=low, high=high, **extra)
        else:
            save(f"ffm_{h}x{w}", img=img, low=low, high=high, fs_re=fs.real, fs_im=fs.imag, **extra)
    g = torch.Generator().manual_seed(9)
    M = torch.rand(36, 64, generator=g)
    save("seploss", M=M, loss=C.Seperation_loss(M))




def gen_fusion():
    """MambaFusion (our GPT-contract wrapper, pure torch glue) built on the REFERENCE's ResidualBlock / MambaConfig:
    the logits oracle of the cross-modal block (SURVEY 7.3 'what the fusion module is')."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from mmidet_b200.mamba import MambaFusion
    torch.manual_seed(1)
    fus = MambaFusion(16, n_layer=2, block_cls=ResidualBlock, config_cls=MambaConfig)
    rgb = torch.randn(2, 16, 6, 5, requires_grad=True)
    ir = torch.randn(2, 16, 6, 5, requires_grad=True)
    o_rgb, o_ir = fus([rgb, ir])
    g1, g2 = torch.randn_like(o_rgb), torch.randn_like(o_ir)
    grgb, gir = torch.autograd.grad([o_rgb, o_ir], [rgb, ir], [g1, g2])
    arrs = {"sd." + k: v for k, v in fus.state_dict().items()}
    save("fusion_block", rgb=rgb, ir=ir, o_rgb=o_rgb, o_ir=o_ir, g_rgb=g1, g_ir=g2, d_rgb=grgb, d_ir=gir, **arrs)


def gen_detector():
    """Two-stream YOLOv5s (BASELINE configs[1], reduced to 160x160 so the fixture stays small): the UNMODIFIED
    models/yolo_test.py Model / parse_model / YAML with `GPT` bound to MambaFusion on reference blocks.  Stored: the
    inputs and outputs of every fusion call site (the boundary our CUDA path must reproduce) and the Detect output
    (for the record; the detector itself cannot travel to the GPU box)."""
    import yaml
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from mmidet_b200.mamba import MambaFusion
    import models.yolo_test as Y
    Y.GPT = lambda d_model, *a, **k: MambaFusion(d_model, n_layer=1, block_cls=ResidualBlock, config_cls=MambaConfig)
    # parse_model compares `m is GPT` after eval()-ing the YAML name, so the bound object must be the same one
    gpt = Y.GPT
    cfg = yaml.safe_load(open(os.path.join(REF, "models/transformer/yolov5l_fusion_transformer_M3FD.yaml")))
    cfg["depth_multiple"], cfg["width_multiple"], cfg["nc"] = 0.33, 0.50, 6
    torch.manual_seed(0)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        model = Y.Model(cfg, ch=3, nc=6).eval()
    fus = [m for m in model.modules() if isinstance(m, MambaFusion)]
    # weights of the fusion blocks are re-drawn from one seed per call site, so the GPU-side test can rebuild them from
    # the seed (same torch build, same constructor draw order) and the fixture only has to carry checksums
    for i, m in enumerate(fus):
        torch.manual_seed(1000 + i)
        m.load_state_dict(MambaFusion(m.n_embd, n_layer=1, block_cls=ResidualBlock, config_cls=MambaConfig).state_dict())
    cap = []
    for m in fus:
        m.register_forward_hook(lambda mod, inp, out: cap.append((inp[0][0].detach(), inp[0][1].detach(), out[0].detach(),
                                                                   out[1].detach())))
    g = torch.Generator().manual_seed(2)
    rgb, ir = torch.rand(1, 3, 96, 96, generator=g), torch.rand(1, 3, 96, 96, generator=g)
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        out = model(rgb, ir)
    det = out[0][0] if isinstance(out[0], (tuple, list)) else out[0]
    arrs = {"detect": det, "n_sites": np.int64(len(fus))}
    for i, (m, (a, b, oa, ob)) in enumerate(zip(fus, cap)):
        arrs.update({f"f{i}.rgb": a, f"f{i}.ir": b, f"f{i}.o_rgb": oa, f"f{i}.o_ir": ob, f"f{i}.d_model": np.int64(m.n_embd)})
        arrs[f"f{i}.wsum"] = np.array([[float(v.double().sum()), float(v.double().abs().sum())]
                                        for _, v in sorted(m.state_dict().items())])
    print("fusion sites:", [tuple(c[0].shape) for c in cap], "detect", tuple(det.shape), "GPT bound:", gpt is Y.GPT)
    save("detector_fusion", **arrs)


def gen_pattern():
    """GPT1_fourier.forward (common.py:357-552) with the transformer stack emptied (n_layer=0: pooling, Fourier split,
    pattern maps, Seperation_loss, conv2 * fea, tokens, pos_emb, ln_f, upsample remain) -- outputs, the token
    embeddings entering self.drop, and gradients of a seeded linear functional of the two output maps."""
    cases = {"pattern_b2": (2, 64, 12, 10, 0, C.GPT1_fourier), "pattern_b9": (9, 16, 9, 11, 1, C.GPT1_fourier),
             "pattern_gpt1": (3, 32, 11, 9, 2, C.GPT1)}  # GPT1 (common.py:140-298): same module without the Fourier branch
    for name, (B, Cc, H, W, seed, cls) in cases.items():
        torch.manual_seed(seed)
        m = cls(Cc, n_layer=0).eval()
        with torch.no_grad():
            m.pos_emb.normal_(0, 0.5)
            m.ln_f.weight.uniform_(0.5, 1.5)
            m.ln_f.bias.normal_(0, 0.1)
            m.conv1.weight.mul_(4.0)  # spread the sigmoids away from 0.5
        g = torch.Generator().manual_seed(100 + seed)
        vis = torch.randn(B, Cc, H, W, generator=g).requires_grad_()
        ir = (torch.randn(B, Cc, H, W, generator=g) * 0.7 + 0.2).requires_grad_()
        g1, g2 = torch.randn(B, Cc, H, W, generator=g), torch.randn(B, Cc, H, W, generator=g)
        cap = {}
        hk = m.drop.register_forward_pre_hook(lambda mod, inp: cap.__setitem__("drop_in", inp[0].detach().clone()))
        pooled = []
        hp = m.avgpool.register_forward_hook(lambda mod, inp, out: pooled.append(out.detach().clone()))
        ro, io, loss = m([vis, ir])
        hk.remove(), hp.remove()
        ((ro * g1).sum() + (io * g2).sum()).backward()
        print(name, "loss", float(loss))
        save(name, vis=vis, ir=ir, g1=g1, g2=g2, conv1_w=m.conv1.weight, conv2_w=m.conv2.weight, pos_emb=m.pos_emb,
             ln_w=m.ln_f.weight, ln_b=m.ln_f.bias, pool_vis=pooled[0], pool_ir=pooled[1], drop_in=cap["drop_in"],
             rgb_out=ro, ir_out=io, loss=loss.detach().reshape(1), d_vis=vis.grad, d_ir=ir.grad,
             d_conv1=m.conv1.weight.grad, d_conv2=m.conv2.weight.grad, d_pos=m.pos_emb.grad)


if __name__ == "__main__":
    torch.set_num_threads(4)
    todo = sys.argv[1:] or ["pscan", "selscan", "block", "step", "ffm", "fusion", "detector", "pattern"]
    for name in todo:
        globals()["gen_" + name]()
