"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy + the C library built from oracle/selscan_oracle.c) of the reference's
fusion hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product (mmidet_b200) never does and has no CPU fallback.

Parity pinning: tests/test_oracle.py checks every function here against tests/golden/*.npz, which are
outputs of the UNMODIFIED reference imported from /root/reference (tests/golden/make_golden.py).

Reference lines restated (all paths relative to the reference root):
  pscan_blelloch / pscan_rev_blelloch   models/pscan.py:37-92, :95-149   (in-place up/down sweeps)
  pscan_forward / pscan_backward        models/pscan.py:152-186, :189-224 (pad to pow2, shift A left, Q)
  selective_scan*                       models/mamba.py:212-233, :235-265, gate :184-186
  extract_frequency2                    models/common.py:37-69 (negative-slice wrap + complex->real->fp16)
  fourier_transform / extract_frequency models/common.py:25-32, :72-93
  separation_loss                       models/common.py:128-139
  rmsnorm / softplus / silu / causal depthwise conv   models/mamba.py:356-366, :203, :180, :126-129,176-178
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile oracle/selscan_oracle.c -> oracle/build/liboracle.so (gcc, OpenMP)."""
    so = os.path.join(_HERE, "build", "liboracle.so")
    src = os.path.join(_HERE, "selscan_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "build", "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


# ----------------------------------------------------------------------------------------------
# pscan: literal numpy restatement of the Blelloch sweeps (models/pscan.py)
# ----------------------------------------------------------------------------------------------
def npo2(n: int) -> int:
    """models/pscan.py:13-18"""
    return 2 ** math.ceil(math.log2(n))


def pad_npo2(X: np.ndarray) -> np.ndarray:
    """models/pscan.py:20-33 -- zero-pad dim 1 of (B, L, D, N) to the next power of two (copies)."""
    Lp = npo2(X.shape[1])
    out = np.zeros((X.shape[0], Lp) + X.shape[2:], dtype=X.dtype)
    out[:, : X.shape[1]] = X
    return out


def pscan_blelloch(A: np.ndarray, X: np.ndarray) -> None:
    """models/pscan.py:37-92.  A, X: (B, D, L, N) with L a power of two; BOTH modified in place."""
    B, D, L, _ = A.shape
    num_steps = int(math.log2(L))
    Aa, Xa = A, X
    for _ in range(num_steps - 2):  # up sweep, :54-63
        Aa = _pairview(Aa)
        Xa = _pairview(Xa)
        Xa[:, :, :, 1] += Aa[:, :, :, 1] * Xa[:, :, :, 0]
        Aa[:, :, :, 1] *= Aa[:, :, :, 0]
        Aa = Aa[:, :, :, 1]
        Xa = Xa[:, :, :, 1]
    if Xa.shape[2] == 4:  # :66-70
        Xa[:, :, 1] += Aa[:, :, 1] * Xa[:, :, 0]
        Aa[:, :, 1] *= Aa[:, :, 0]
        Xa[:, :, 3] += Aa[:, :, 3] * (Xa[:, :, 2] + Aa[:, :, 2] * Xa[:, :, 1])
    elif Xa.shape[2] == 2:  # :71-73
        Xa[:, :, 1] += Aa[:, :, 1] * Xa[:, :, 0]
        return
    else:
        return
    s = 2 ** (num_steps - 2)  # down sweep, :78-92
    Aa = A[:, :, s - 1 : L : s]
    Xa = X[:, :, s - 1 : L : s]
    Xa[:, :, 2] += Aa[:, :, 2] * Xa[:, :, 1]
    Aa[:, :, 2] *= Aa[:, :, 1]
    for k in range(num_steps - 3, -1, -1):
        s = 2**k
        Aa = _pairview(A[:, :, s - 1 : L : s])
        Xa = _pairview(X[:, :, s - 1 : L : s])
        Xa[:, :, 1:, 0] += Aa[:, :, 1:, 0] * Xa[:, :, :-1, 1]
        Aa[:, :, 1:, 0] *= Aa[:, :, :-1, 1]


def _pairview(a: np.ndarray) -> np.ndarray:
    """view (B, D, T, N) -> (B, D, T/2, 2, N) without copying (torch .view on a strided slice)."""
    B, D, T, N = a.shape
    s = a.strides
    return np.lib.stride_tricks.as_strided(a, shape=(B, D, T // 2, 2, N), strides=(s[0], s[1], 2 * s[2], s[2], s[3]))


def pscan_rev_blelloch(A: np.ndarray, X: np.ndarray) -> None:
    """models/pscan.py:95-149 -- the same sweeps reversed in time; in place."""
    B, D, L, _ = A.shape
    num_steps = int(math.log2(L))
    Aa, Xa = A, X
    for _ in range(num_steps - 2):
        Aa = _pairview(Aa)
        Xa = _pairview(Xa)
        Xa[:, :, :, 0] += Aa[:, :, :, 0] * Xa[:, :, :, 1]
        Aa[:, :, :, 0] *= Aa[:, :, :, 1]
        Aa = Aa[:, :, :, 0]
        Xa = Xa[:, :, :, 0]
    if Xa.shape[2] == 4:
        Xa[:, :, 2] += Aa[:, :, 2] * Xa[:, :, 3]
        Aa[:, :, 2] *= Aa[:, :, 3]
        Xa[:, :, 0] += Aa[:, :, 0] * (Xa[:, :, 1] + Aa[:, :, 1] * Xa[:, :, 2])
    elif Xa.shape[2] == 2:
        Xa[:, :, 0] += Aa[:, :, 0] * Xa[:, :, 1]
        return
    else:
        return
    s = 2 ** (num_steps - 2)
    Aa = A[:, :, 0:L:s]
    Xa = X[:, :, 0:L:s]
    Xa[:, :, 1] += Aa[:, :, 1] * Xa[:, :, 2]
    Aa[:, :, 1] *= Aa[:, :, 2]
    for k in range(num_steps - 3, -1, -1):
        s = 2**k
        Aa = _pairview(A[:, :, 0:L:s])
        Xa = _pairview(X[:, :, 0:L:s])
        Xa[:, :, :-1, 1] += Aa[:, :, :-1, 1] * Xa[:, :, 1:, 0]
        Aa[:, :, :-1, 1] *= Aa[:, :, 1:, 0]


def pscan_forward(A_in: np.ndarray, X_in: np.ndarray):
    """models/pscan.py:152-186.  (B, L, D, N) -> H (B, L, D, N); also returns the padded (B, D, Lp, N) H that
    the reference saves for backward."""
    L = X_in.shape[1]
    A = pad_npo2(A_in) if L != npo2(L) else A_in.copy()
    X = pad_npo2(X_in) if L != npo2(L) else X_in.copy()
    A = np.ascontiguousarray(A.transpose(0, 2, 1, 3))
    X = np.ascontiguousarray(X.transpose(0, 2, 1, 3))
    pscan_blelloch(A, X)
    return X.transpose(0, 2, 1, 3)[:, :L], X


def pscan_backward(A_in: np.ndarray, Hpad: np.ndarray, grad_in: np.ndarray):
    """models/pscan.py:189-224 -> (gradA, gradX), both (B, L, D, N)."""
    L = grad_in.shape[1]
    g = pad_npo2(grad_in) if L != npo2(L) else grad_in.copy()
    Ap = pad_npo2(A_in) if L != npo2(L) else A_in
    g = np.ascontiguousarray(g.transpose(0, 2, 1, 3))
    Ap = Ap.transpose(0, 2, 1, 3)
    A = np.zeros_like(g)
    A[:, :, :-1] = Ap[:, :, 1:]  # :216 shift one step left, zero at the end
    pscan_rev_blelloch(A, g)
    Q = np.zeros_like(Hpad)
    Q[:, :, 1:] += Hpad[:, :, :-1] * g[:, :, 1:]  # :221-222
    return Q.transpose(0, 2, 1, 3)[:, :L], g.transpose(0, 2, 1, 3)[:, :L]


# sequential (C) statement of the same recurrences ------------------------------------------------
def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32", np.float32
    if dtype == np.float64:
        return "f64", np.float64
    raise TypeError(dtype)


def pscan_seq_fwd(A, X):
    sfx, dt = _sfx(A.dtype)
    A, X = _c(A, dt), _c(X, dt)
    H = np.empty_like(X)
    B, L, D, N = A.shape
    getattr(lib(), f"oracle_pscan_fwd_{sfx}")(_p(A), _p(X), _p(H), B, L, D, N)
    return H


def pscan_seq_bwd(A, H, gH):
    sfx, dt = _sfx(A.dtype)
    A, H, gH = _c(A, dt), _c(H, dt), _c(gH, dt)
    gA, gX = np.empty_like(A), np.empty_like(A)
    B, L, D, N = A.shape
    getattr(lib(), f"oracle_pscan_bwd_{sfx}")(_p(A), _p(H), _p(gH), _p(gA), _p(gX), B, L, D, N)
    return gA, gX


# ----------------------------------------------------------------------------------------------
# selective scan (models/mamba.py:212-265) with optional SiLU gate (:184-186)
# ----------------------------------------------------------------------------------------------
def selective_scan_fwd(x, delta, A, Bm, Cm, D, z=None, h0=None, dtype=np.float32, return_state=False):
    sfx, dt = _sfx(dtype)
    x, delta, A, Bm, Cm, D, z, h0 = (_c(v, dt) for v in (x, delta, A, Bm, Cm, D, z, h0))
    B, L, ED = x.shape
    N = A.shape[1]
    assert N <= 64
    out = np.empty_like(x)
    hT = np.empty((B, ED, N), dt) if return_state else None
    getattr(lib(), f"oracle_selscan_fwd_{sfx}")(
        _p(x), _p(delta), _p(z), _p(A), _p(Bm), _p(Cm), _p(D), _p(h0), _p(out), _p(hT), B, L, ED, N)
    return (out, hT) if return_state else out


def selective_scan_bwd(x, delta, A, Bm, Cm, D, dout, z=None, dtype=np.float32):
    """-> dict(dx, ddelta, dz, dA, dB, dC, dD)"""
    sfx, dt = _sfx(dtype)
    x, delta, A, Bm, Cm, D, z, dout = (_c(v, dt) for v in (x, delta, A, Bm, Cm, D, z, dout))
    B, L, ED = x.shape
    N = A.shape[1]
    dx, dd = np.empty_like(x), np.empty_like(x)
    dz = np.empty_like(x) if z is not None else None
    dA, dB, dC, dD = np.empty_like(A), np.empty_like(Bm), np.empty_like(Cm), np.empty_like(D)
    getattr(lib(), f"oracle_selscan_bwd_{sfx}")(
        _p(x), _p(delta), _p(z), _p(A), _p(Bm), _p(Cm), _p(D), _p(dout), _p(dx), _p(dd), _p(dz), _p(dA), _p(dB),
        _p(dC), _p(dD), B, L, ED, N)
    return dict(dx=dx, ddelta=dd, dz=dz, dA=dA, dB=dB, dC=dC, dD=dD)


def selective_scan_pscan(x, delta, A, Bm, Cm, D):
    """models/mamba.py:212-233 literally: materialise deltaA/BX, Blelloch pscan, readout."""
    deltaA = np.exp(delta[..., None] * A)
    BX = (delta[..., None] * Bm[:, :, None, :]) * x[..., None]
    hs, _ = pscan_forward(deltaA, BX)
    y = np.einsum("bldn,bln->bld", hs, Cm)  # (hs @ C.unsqueeze(-1)).squeeze(3), mamba.py:229
    return y + D * x


# ----------------------------------------------------------------------------------------------
# the rest of MambaBlock around the scan (models/mamba.py) -- numpy, for module-level goldens
# ----------------------------------------------------------------------------------------------
def silu(v):
    return v / (1.0 + np.exp(-v))


def softplus(v):
    """F.softplus, beta=1, threshold=20 (models/mamba.py:203)"""
    return np.where(v > 20.0, v, np.log1p(np.exp(np.minimum(v, 20.0))))


def rmsnorm(x, w, eps=1e-5):
    """models/mamba.py:356-366"""
    return x * (1.0 / np.sqrt(np.mean(x * x, axis=-1, keepdims=True) + eps)) * w


def causal_dwconv(x, w, b):
    """models/mamba.py:126-129,176-178: depthwise Conv1d(k, padding=k-1)[:, :, :L] on (B, L, ED)."""
    B, L, ED = x.shape
    k = w.shape[-1]
    w = w.reshape(ED, k)
    xp = np.concatenate([np.zeros((B, k - 1, ED), x.dtype), x], axis=1)
    y = np.zeros_like(x)
    for j in range(k):
        y += xp[:, j : j + L] * w[:, j]
    return y + (0 if b is None else b)


def mamba_block_forward(x, p, d_state=16, dt_rank=None, dtype=np.float32):
    """models/mamba.py:165-210 with a reference state_dict `p` (numpy arrays keyed like MambaBlock)."""
    x = x.astype(dtype)
    ED = p["A_log"].shape[0]
    dt_rank = dt_rank or p["dt_proj.weight"].shape[1]
    xz = x @ p["in_proj.weight"].T.astype(dtype)
    xs, z = xz[..., :ED], xz[..., ED:]
    xs = silu(causal_dwconv(xs, p["conv1d.weight"].astype(dtype), p["conv1d.bias"].astype(dtype)))
    A = -np.exp(p["A_log"].astype(dtype))
    dbc = xs @ p["x_proj.weight"].T.astype(dtype)
    dl, Bm, Cm = dbc[..., :dt_rank], dbc[..., dt_rank : dt_rank + d_state], dbc[..., dt_rank + d_state :]
    delta = softplus(dl @ p["dt_proj.weight"].T.astype(dtype) + p["dt_proj.bias"].astype(dtype))
    y = selective_scan_fwd(xs, delta, A, Bm, Cm, p["D"], z=z, dtype=dtype)
    return y @ p["out_proj.weight"].T.astype(dtype)


# ----------------------------------------------------------------------------------------------
# Fusion Focus Module Fourier step (models/common.py)
# ----------------------------------------------------------------------------------------------
def fourier_transform(image):
    """models/common.py:25-32"""
    return np.fft.fftshift(np.fft.fftn(image, axes=(-2, -1)), axes=(-2, -1))


def ffm_masks(rows: int, cols: int):
    """Boolean masks over the SHIFTED spectrum equivalent to the slice assignments of
    models/common.py:44-56, including Python's negative-index wrap (threshold > crow).
    Returns (keep_high, keep_low): high-pass keeps where keep_high, low-pass keeps where keep_low."""
    crow, ccol = rows // 2, cols // 2
    thr = crow + ccol // 4
    rsel = np.zeros(rows, bool)
    rsel[slice(crow - thr, crow + thr)] = True  # rows zeroed in the high-pass (:47-48)
    csel = np.zeros(cols, bool)
    csel[slice(ccol - thr, ccol + thr)] = True
    keep_high = ~np.outer(rsel, csel)
    rkeep = np.ones(rows, bool)
    rkeep[slice(None, crow - thr)] = False  # :51
    rkeep[slice(crow + thr, None)] = False  # :52
    ckeep = np.ones(cols, bool)
    ckeep[slice(None, ccol - thr)] = False  # :53
    ckeep[slice(ccol + thr, None)] = False  # :54
    keep_low = np.outer(rkeep, ckeep)
    return keep_high, keep_low


def extract_frequency2(image):
    """models/common.py:37-69 -> (low, high), both float16 real (B, C, H, W)."""
    image = np.asarray(image)
    f = np.fft.fftn(image.astype(np.float32), axes=(-2, -1)).astype(np.complex64)
    fs = np.fft.fftshift(f, axes=(-2, -1))
    rows, cols = image.shape[-2:]
    crow, ccol = rows // 2, cols // 2
    thr = crow + ccol // 4
    hp = fs.copy()
    hp[:, :, crow - thr : crow + thr, ccol - thr : ccol + thr] = 0
    lp = fs.copy()
    lp[:, :, : crow - thr, :] = 0
    lp[:, :, crow + thr :, :] = 0
    lp[:, :, :, : ccol - thr] = 0
    lp[:, :, :, ccol + thr :] = 0
    high = np.fft.ifftn(np.fft.ifftshift(hp, axes=(-2, -1)), axes=(-2, -1))
    low = np.fft.ifftn(np.fft.ifftshift(lp, axes=(-2, -1)), axes=(-2, -1))
    # complex -> .half() keeps the real part (common.py:66-67)
    return low.real.astype(np.float32).astype(np.float16), high.real.astype(np.float32).astype(np.float16)


def extract_frequency(image, threshold=30):
    """models/common.py:72-93 (dead code in the reference; restated for completeness)."""
    fs = fourier_transform(np.asarray(image, np.float32)).astype(np.complex64)
    H, W = image.shape[-2:]
    ch, cw = H // 2, W // 2
    low = fs.copy()
    low[:, :, ch - threshold : ch + threshold, cw - threshold : cw + threshold] = 0
    high = fs - low
    return low.real.astype(np.float16), high.real.astype(np.float16)


def separation_loss(M):
    """models/common.py:128-139: (sum_{i<j} M_i . M_j) / (l (l-1)); rows of M are 1-D patterns."""
    M = np.asarray(M, np.float64)
    l = M.shape[0]
    acc = 0.0
    for i in range(l - 1):
        for j in range(i + 1, l):
            acc += float(M[i] @ M[j])
    return acc / (l * (l - 1))


# ----------------------------------------------------------------------------------------------
# depthwise causal conv1d + SiLU on channels-last tokens (models/mamba.py:176-180 with nn.Conv1d of :125-128)
def adaptive_avg_pool2d(x, size):
    """nn.AdaptiveAvgPool2d(size) as GPT1_fourier uses it (models/common.py:324-325, :396-397): cell (i, j) averages
    rows [floor(i H / hs), ceil((i + 1) H / hs)) and the matching columns (windows overlap when H % hs != 0)."""
    x = np.asarray(x, np.float64)
    H, W = x.shape[-2:]
    hs, ws = size
    out = np.empty(x.shape[:-2] + (hs, ws))
    for i in range(hs):
        r0, r1 = (i * H) // hs, -((-(i + 1) * H) // hs)
        for j in range(ws):
            c0, c1 = (j * W) // ws, -((-(j + 1) * W) // ws)
            out[..., i, j] = x[..., r0:r1, c0:c1].mean(axis=(-2, -1))
    return out


def _bilinear_taps(n_out, n_in):
    """align_corners=False source taps of F.interpolate(mode='bilinear') (models/common.py:540-543)."""
    src = np.maximum((np.arange(n_out) + 0.5) * (n_in / n_out) - 0.5, 0.0)
    i0 = np.minimum(src.astype(np.int64), n_in - 1)
    i1 = np.minimum(i0 + 1, n_in - 1)
    lam = src - i0
    return i0, i1, lam


def upsample_bilinear(x, size):
    """F.interpolate(x, size=size, mode='bilinear') with the default align_corners=False."""
    x = np.asarray(x, np.float64)
    y0, y1, ly = _bilinear_taps(size[0], x.shape[-2])
    x0, x1, lx = _bilinear_taps(size[1], x.shape[-1])
    top = x[..., y0, :][..., :, x0] * (1 - lx) + x[..., y0, :][..., :, x1] * lx
    bot = x[..., y1, :][..., :, x0] * (1 - lx) + x[..., y1, :][..., :, x1] * lx
    return top * (1 - ly)[:, None] + bot * ly[:, None]


def ffm_pattern(pool_vis, pool_ir, conv1_w, conv2_w, high=True):
    """models/common.py:434-516 (GPT1_fourier.forward between avgpool and the transformer), literally
    (high=False: GPT1.forward, models/common.py:218-262 -- no Fourier branch, loss over [M_vis; M_ir]):
    pooled maps (B, C, h, w), conv1_w (8, C), conv2_w (C, 8) -> (token_embeddings (B, 2hw, C), pattenLoss).
    1x1 convolutions are einsums over the channel axis; `.view(-1, h*w)` flattens (b, j) row-major."""
    B, Cc, h, w = pool_vis.shape
    sig = lambda v: 1.0 / (1.0 + np.exp(-v))  # noqa: E731
    conv1 = lambda t: np.einsum("jc,bchw->bjhw", conv1_w.astype(np.float64), t.astype(np.float64))  # noqa: E731
    rows_high, rows, toks = [], [], []
    for fea in (pool_vis, pool_ir):
        _, hi_pass = extract_frequency2(fea)                                # :434-435
        high_multi = hi_pass.astype(np.float32) * fea                       # :440-441 (fp16 * fp32 -> fp32)
        rows_high.append(sig(conv1(high_multi)).reshape(-1, h * w))         # :444-455
        M = sig(conv1(fea))                                                 # :476-480
        rows.append(M.reshape(-1, h * w))                                   # :482-483
        PT = np.einsum("cj,bjhw->bchw", conv2_w.astype(np.float64), M)      # :496-497
        toks.append((PT * fea).reshape(B, Cc, -1))                          # :499-503
    n_half = len(rows_high[0]) // 8                                         # :487
    cat = np.concatenate([rows[0], rows[1], rows_high[0][:n_half], rows_high[1][:n_half]], axis=0)  # :488-489
    if not high:
        cat = np.concatenate([rows[0], rows[1]], axis=0)                    # GPT1: models/common.py:233-235
    loss = separation_loss(cat)                                             # :494
    tok = np.concatenate(toks, axis=2).transpose(0, 2, 1)                   # :514-519
    return np.ascontiguousarray(tok), loss


def causal_conv1d_silu(x, w, bias=None, silu=True, dtype=np.float64):
    """x (B, L, ED); w (ED, K) [= conv1d.weight[:, 0, :]]; pre[t] = bias + sum_j w[:, j] * x[t-(K-1)+j]; y = silu(pre)."""
    x = np.asarray(x, dtype)
    w = np.asarray(w, dtype)
    Bsz, L, ED = x.shape
    K = w.shape[1]
    xp = np.concatenate([np.zeros((Bsz, K - 1, ED), dtype), x], axis=1)
    pre = np.zeros((Bsz, L, ED), dtype) + (0 if bias is None else np.asarray(bias, dtype))
    for j in range(K):
        pre = pre + w[:, j] * xp[:, j:j + L]
    return pre / (1 + np.exp(-pre)) if silu else pre


def causal_conv1d_silu_bwd(x, w, bias, dy, silu=True, dtype=np.float64):
    """-> dx (B, L, ED), dw (ED, K), dbias (ED): adjoint of causal_conv1d_silu."""
    x = np.asarray(x, dtype)
    w = np.asarray(w, dtype)
    dy = np.asarray(dy, dtype)
    Bsz, L, ED = x.shape
    K = w.shape[1]
    xp = np.concatenate([np.zeros((Bsz, K - 1, ED), dtype), x], axis=1)
    pre = np.zeros((Bsz, L, ED), dtype) + (0 if bias is None else np.asarray(bias, dtype))
    for j in range(K):
        pre = pre + w[:, j] * xp[:, j:j + L]
    if silu:
        s = 1 / (1 + np.exp(-pre))
        dpre = dy * s * (1 + pre * (1 - s))
    else:
        dpre = dy
    dxp = np.zeros_like(xp)
    dw = np.zeros_like(w)
    for j in range(K):
        dxp[:, j:j + L] += w[:, j] * dpre
        dw[:, j] = (dpre * xp[:, j:j + L]).sum((0, 1))
    return dxp[:, K - 1:], dw, dpre.sum((0, 1))
