#!/usr/bin/env python
"""Timings of the parity-API / FFM kernels (not the headline path): pscan fwd/bwd on materialised tensors, the FFM
Fourier step at the reference's 8x8 pooled size, the closed-form separation loss, the channels-last conv prologue."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmidet_b200 import ops
from mmidet_b200.ffm import extract_frequency2, separation_loss
from mmidet_b200.pscan import pscan

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

dev = "cuda"
B, L, D, N = 2, 6400, 256, 16
A = (torch.rand(B, L, D, N, device=dev) * 0.9 + 0.05).requires_grad_(True)
X = torch.randn(B, L, D, N, device=dev, requires_grad=True)
gH = torch.randn(B, L, D, N, device=dev)
nbytes = B * L * D * N * 4
tf = timeit(lambda: pscan(A, X))
H = pscan(A, X)
tb = timeit(lambda: torch.autograd.grad(H, (A, X), gH, retain_graph=True))
print(f"pscan (B={B},L={L},D={D},N={N}) fwd {tf:.3f} ms = {3*nbytes/tf/1e6:.0f} GB/s algorithmic (3 tensor passes); "
      f"bwd {tb:.3f} ms = {5*nbytes/tb/1e6:.0f} GB/s (5 passes)")
for bc in (2 * 128, 16 * 128):
    img = torch.randn(bc // 128, 128, 8, 8, device=dev)
    t = timeit(lambda: extract_frequency2(img, with_product=True), 50)
    print(f"ffm extract_frequency2 on ({bc//128},128,8,8): {t*1e3:.1f} us per call (one launch; reference: ~18 launches per modality)")
M = torch.rand(288, 64, device=dev)
t = timeit(lambda: separation_loss(M), 50)
print(f"separation_loss l=288 (B=16): {t*1e3:.1f} us (reference: O(l^2) python loop, 615 ms on CPU)")
Bc, Lc, ED = 16, 6400, 512
xz = torch.randn(Bc, Lc, 2 * ED, device=dev)
w = torch.randn(ED, 1, 4, device=dev); b = torch.randn(ED, device=dev)
x = xz.chunk(2, -1)[0]
t = timeit(lambda: ops.causal_conv1d_silu(x, w, b), 20)
byt = Bc * Lc * ED * 4 * 2
print(f"causal conv1d+SiLU fwd (B={Bc},L={Lc},ED={ED}) {t:.3f} ms = {byt/t/1e6:.0f} GB/s (read x + write y)")
conv = torch.nn.Conv1d(ED, ED, 4, groups=ED, padding=3).cuda()
t2 = timeit(lambda: torch.nn.functional.silu(conv(x.transpose(1, 2))[:, :, :Lc].transpose(1, 2)).contiguous(), 10)
print(f"  same op via transpose + nn.Conv1d + transpose + silu (stock PyTorch): {t2:.3f} ms")

# ---- the remaining "next"-row kernels at the training shape of the P3 fusion block (16 pairs, 80x80 maps, d_model 256):
# algorithmic bytes / time vs the measured HBM peak.  Working sets are >= 200 MB (> L2), timed back to back.
PEAK = 6538.0


def report(name, ms, nbytes, extra=""):
    print(f"{name}: {ms:.3f} ms = {nbytes / ms / 1e6:.0f} GB/s = {nbytes / ms / 1e6 / PEAK * 100:.0f} % of the HBM peak{extra}")


for dt in (torch.float32, torch.bfloat16):
    es = torch.empty(0, dtype=dt).element_size()
    tag = "fp32" if dt == torch.float32 else "bf16"
    Bt, C, Hm, Wm = 16, 256, 80, 80
    Lt, EDt = 2 * Hm * Wm, 2 * C
    # causal conv backward: read x, dy; write dx (+ dw, dbias)
    xc = torch.randn(Bt, Lt, EDt, device=dev, dtype=dt, requires_grad=True)
    wc = torch.randn(EDt, 1, 4, device=dev, requires_grad=True); bcv = torch.randn(EDt, device=dev, requires_grad=True)
    yc = ops.causal_conv1d_silu(xc, wc, bcv)
    gy = torch.randn_like(yc)
    t = timeit(lambda: torch.autograd.grad(yc, (xc, wc, bcv), gy, retain_graph=True), 20)
    report(f"causal conv1d+SiLU bwd {tag} (B={Bt},L={Lt},ED={EDt})", t, 3 * Bt * Lt * EDt * es)
    t = timeit(lambda: ops.causal_conv1d_silu(xc, wc, bcv), 20)
    report(f"causal conv1d+SiLU fwd {tag} (same shape)", t, 2 * Bt * Lt * EDt * es)
    # RMSNorm on the token stream (B, L, d_model)
    xr = torch.randn(Bt, Lt, C, device=dev, dtype=dt, requires_grad=True)
    wr = torch.ones(C, device=dev, requires_grad=True)
    yr = ops.rmsnorm(xr, wr)
    gr = torch.randn_like(yr)
    t = timeit(lambda: ops.rmsnorm(xr, wr), 20)
    report(f"RMSNorm fwd {tag} (B={Bt},L={Lt},C={C})", t, 2 * Bt * Lt * C * es)
    t = timeit(lambda: torch.autograd.grad(yr, (xr, wr), gr, retain_graph=True), 20)
    report(f"RMSNorm bwd {tag}", t, 3 * Bt * Lt * C * es)
    # token layout: two NCHW maps <-> (B, 2HW, C) tokens
    rgb = torch.randn(Bt, C, Hm, Wm, device=dev, dtype=dt); ir = torch.randn_like(rgb)
    tok = ops.tokens_gather(rgb, ir)
    t = timeit(lambda: ops.tokens_gather(rgb, ir), 20)
    report(f"tokens gather {tag} (2 x ({Bt},{C},{Hm},{Wm}) -> ({Bt},{Lt},{C}))", t, 2 * tok.numel() * es)
    t2 = timeit(lambda: torch.cat([rgb.flatten(2), ir.flatten(2)], dim=2).transpose(1, 2).contiguous(), 20)
    print(f"  same layout change in stock torch (flatten / cat / transpose / contiguous): {t2:.3f} ms")
    t = timeit(lambda: ops.tokens_scatter(tok, rgb.shape), 20)
    report(f"tokens scatter {tag}", t, 2 * tok.numel() * es)

    # kernel-only times of the two backward kernels (direct C-ABI calls on preallocated buffers: the autograd path above
    # adds 60-100 us of host work per call, which hides kernels this short)
    from mmidet_b200 import _lib
    lib = _lib.load()
    P, DT, ST = ops._ptr, ops._DT, ops._stream
    x2, g2 = xr.detach().reshape(-1, C), gr.reshape(-1, C)
    dx2, dw2 = torch.empty_like(x2), torch.empty(C, device=dev)
    w32 = wr.detach().float()
    t = timeit(lambda: lib.mmi_rmsnorm_bwd(P(x2), P(w32), P(g2), P(dx2), P(dw2), x2.shape[0], C, x2.stride(0), g2.stride(0),
                                           dx2.stride(0), 1e-5, DT[dt], -1, ST(x2)), 20)
    report(f"RMSNorm bwd kernel only {tag}", t, 3 * Bt * Lt * C * es)
    xc2, gy2 = xc.detach(), gy
    dxc, dwc, dbc = torch.empty_like(xc2), torch.empty(EDt, 4, device=dev), torch.empty(EDt, device=dev)
    wc2, bc2 = wc.detach().reshape(EDt, 4).contiguous(), bcv.detach()
    t = timeit(lambda: lib.mmi_causal_conv1d_bwd(P(xc2), P(wc2), P(bc2), P(gy2), P(dxc), P(dwc), P(dbc), Bt, Lt, EDt, 4, xc2.stride(1),
                                                 gy2.stride(1), dxc.stride(1), DT[dt], 1, ST(xc2)), 20)
    report(f"causal conv1d+SiLU bwd kernel only {tag}", t, 3 * Bt * Lt * EDt * es)
    yc2 = torch.empty_like(xc2)
    t = timeit(lambda: lib.mmi_causal_conv1d_fwd(P(xc2), P(wc2), P(bc2), P(yc2), Bt, Lt, EDt, 4, xc2.stride(1), yc2.stride(1), DT[dt], 1,
                                                 ST(xc2)), 20)
    report(f"causal conv1d+SiLU fwd kernel only {tag}", t, 2 * Bt * Lt * EDt * es)
