#!/usr/bin/env python
"""scripts/ffm_bench.py -- pattern path of the Fusion Focus Module (models/common.py:434-516) on one B200:
the fused kernels (ffm.pattern_tokens: Fourier split of the rows that reach the loss + one pattern launch + loss)
against the same op sequence the reference issues, restated with stock torch ops on the GPU (torch.fft, 1x1 conv2d,
sigmoid, cat/permute; the separation loss in closed form AND as the reference's O(l^2) Python loop).
CUDA-event timed, median of 50 after 10 warm-ups.  Prints one line per batch size."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmidet_b200 import ffm, ops  # noqa: E402


def torch_extract(image):
    """extract_frequency2 with torch.fft and the kept block of ffm.kept_range (same masks as common.py:41-56)."""
    H, W = image.shape[-2:]
    r0, r1, c0, c1 = ffm.kept_range(H, W)
    fs = torch.fft.fftshift(torch.fft.fftn(image.float(), dim=(-2, -1)), dim=(-2, -1))  # torch.fft has no bf16
    lo, hi = torch.zeros_like(fs), fs.clone()
    lo[..., r0:r1, c0:c1] = fs[..., r0:r1, c0:c1]
    hi[..., r0:r1, c0:c1] = 0
    inv = lambda t: torch.fft.ifftn(torch.fft.ifftshift(t, dim=(-2, -1)), dim=(-2, -1)).real.half()  # noqa: E731
    return inv(lo), inv(hi)


def sep_closed(M):
    l = M.shape[0]
    return ((M.sum(0) ** 2).sum() - (M * M).sum()) / (2 * l * (l - 1))


def sep_loop(M):  # common.py:128-139 as written
    l = len(M)
    loss = 0
    for i in range(l - 1):
        for j in range(i + 1, l):
            loss = loss + torch.dot(M[i], M[j])
    return loss / (l * (l - 1))


def torch_pattern(vis, ir, conv1, conv2, sep):
    rows, highs, toks = [], [], []
    for fea in (vis, ir):
        _, high = torch_extract(fea)
        highs.append(torch.sigmoid(conv1(high * fea)).view(-1, 64))
        M = torch.sigmoid(conv1(fea))
        rows.append(M.view(-1, 64))
        toks.append((conv2(M) * fea).flatten(2))
    n = len(highs[0]) // 8
    loss = sep(torch.cat((rows[0], rows[1], highs[0][:n], highs[1][:n]), dim=0))
    return torch.cat(toks, dim=2).permute(0, 2, 1).contiguous(), loss


def timed(fn, iters=50, warm=10):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    torch.manual_seed(0)
    C = 128
    conv1 = torch.nn.Conv2d(C, 8, 1, bias=False).cuda()
    conv2 = torch.nn.Conv2d(8, C, 1, bias=False).cuda()
    for B in (2, 16, 64):
        vis = torch.randn(B, C, 8, 8, device="cuda", requires_grad=True)
        ir = torch.randn(B, C, 8, 8, device="cuda", requires_grad=True)
        dtok = torch.randn(B, 128, C, device="cuda")

        def ours(bwd):
            tok, loss = ffm.pattern_tokens(vis, ir, conv1.weight, conv2.weight)
            if bwd:
                torch.autograd.grad(tok, (vis, ir, conv1.weight, conv2.weight), dtok)

        def stock(bwd, sep):
            tok, loss = torch_pattern(vis, ir, conv1, conv2, sep)
            if bwd:
                torch.autograd.grad(tok, (vis, ir, conv1.weight, conv2.weight), dtok)

        n0 = ops.launches
        ours(True)
        row = {"B": B, "C": C, "P": 64, "rows_in_loss": 18 * B, "our_launches_fwd_bwd": ops.launches - n0,
               "ours_fwd_ms": round(timed(lambda: ours(False)), 4), "ours_fwd_bwd_ms": round(timed(lambda: ours(True)), 4),
               "torch_fwd_ms": round(timed(lambda: stock(False, sep_closed)), 4),
               "torch_fwd_bwd_ms": round(timed(lambda: stock(True, sep_closed)), 4)}
        if B <= 2:
            row["torch_fwd_loop_loss_ms"] = round(timed(lambda: stock(False, sep_loop), iters=5, warm=1), 3)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
