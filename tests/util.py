"""Shared helpers for the parity tests: seeded synthetic inputs and error metrics."""
import numpy as np


def relerr(a, b):
    """max |a-b| / max |b|  (scaled max error; the tolerance of north_star is stated on this)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def scan_inputs(B, L, ED, N=16, seed=0, random_A=False, small_delta=False):
    """x~N(0,1), delta=softplus(N(0,1)-3), B,C~N(0,1), A=-exp(A_log) (S4D-real init or random), D=1 or random."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, L, ED)).astype(np.float32)
    delta = np.log1p(np.exp(rng.standard_normal((B, L, ED)) - (6.0 if small_delta else 3.0))).astype(np.float32)
    z = rng.standard_normal((B, L, ED)).astype(np.float32)
    Bm = rng.standard_normal((B, L, N)).astype(np.float32)
    Cm = rng.standard_normal((B, L, N)).astype(np.float32)
    if random_A:
        A = (-np.exp(rng.standard_normal((ED, N)) * 0.7 + 0.5)).astype(np.float32)
        D = rng.standard_normal(ED).astype(np.float32)
    else:
        A = (-np.exp(np.log(np.tile(np.arange(1, N + 1, dtype=np.float32), (ED, 1))))).astype(np.float32)
        D = np.ones(ED, np.float32)
    dout = rng.standard_normal((B, L, ED)).astype(np.float32)
    return dict(x=x, delta=delta, z=z, A=A, Bm=Bm, Cm=Cm, D=D, dout=dout)
