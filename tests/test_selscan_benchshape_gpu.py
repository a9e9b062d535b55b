"""GPU parity AT THE BENCHMARKED CONFIGURATIONS: values (not properties) of the fused selective scan against the
fp64 C oracle, at the shapes bench.py, scripts/sweep.py and BASELINE.json's configs name.

The kernels pick a different CTA schedule per shape (L split / chained L segments / persistent grid), so the small-shape
tests of test_selscan_gpu.py do not cover the schedule the benchmark runs: here the exact benchmark shapes are run and
every output is compared, one batch entry at a time (the oracle costs ~0.2 s per B=1 entry at L=6400, ED=512).

Tolerances (north_star): rel-err max|a-b|/max|b| <= 1e-4 for fp32 I/O, <= 2e-2 for bf16 I/O (fp16: 5e-3)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from tests.util import relerr

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2, torch.float16: 5e-3}


def _inputs(B, L, ED, dtype, random_A, seed, N=16):
    """bench.py's generator (torch, on the device): x,z,B,C,dout ~ N(0,1), delta = softplus(N(0,1) - 3)."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
    x, z, dout = rn(B, L, ED), rn(B, L, ED), rn(B, L, ED)
    delta = torch.nn.functional.softplus(rn(B, L, ED) - 3.0)
    Bm, Cm = rn(B, L, N), rn(B, L, N)
    if random_A:  # a trained A_log: every row leaves the geometric fast path
        A = -torch.exp(rn(ED, N) * 0.5 + 0.3)
        D = rn(ED)
    else:  # S4D-real init, models/mamba.py:158-159
        A = -torch.arange(1, N + 1, device="cuda", dtype=torch.float32).repeat(ED, 1)
        D = torch.ones(ED, device="cuda")
    t = {k: v.to(dtype) for k, v in dict(x=x, delta=delta, z=z, Bm=Bm, Cm=Cm, dout=dout).items()}
    t.update(A=A.contiguous(), D=D)
    return t


def _check(B, L, ED, dtype=torch.float32, random_A=False, flags=0, entries=None, seed=0, gate=True):
    from mmidet_b200 import ops
    t = _inputs(B, L, ED, dtype, random_A, seed)
    z = t["z"] if gate else None
    out, _, chk, saved = ops.selscan_fwd_raw(t["x"], t["delta"], t["A"], t["Bm"], t["Cm"], t["D"], z=z, want_chk=True,
                                             flags=flags)
    dx, dd, dz, dA, dB, dC, dD = ops.selscan_bwd_raw(saved, chk, t["dout"], flags=flags)
    torch.cuda.synchronize()
    tol = TOL[dtype]
    f64 = lambda v: v.detach().double().cpu().numpy()
    entries = list(range(B)) if entries is None else entries
    A64, D64 = f64(t["A"]), f64(t["D"])
    dA_ref, dD_ref = np.zeros_like(A64), np.zeros_like(D64)
    bad = {}
    for b in range(B):
        if b not in entries:
            continue
        s = slice(b, b + 1)
        a = {k: f64(t[k][s]) for k in ("x", "delta", "z", "Bm", "Cm", "dout")}
        ref_out = O.selective_scan_fwd(a["x"], a["delta"], A64, a["Bm"], a["Cm"], D64, z=a["z"] if gate else None,
                                       dtype=np.float64)
        ref = O.selective_scan_bwd(a["x"], a["delta"], A64, a["Bm"], a["Cm"], D64, a["dout"],
                                   z=a["z"] if gate else None, dtype=np.float64)
        dA_ref += ref["dA"]
        dD_ref += ref["dD"]
        got = dict(out=out[s], dx=dx[s], ddelta=dd[s], dB=dB[s], dC=dC[s])
        ref["out"] = ref_out
        if gate:
            got["dz"] = dz[s]
        for k, v in got.items():
            e = relerr(f64(v), ref[k])
            if not (e <= tol):
                bad[f"{k}[b={b}]"] = e
    if len(entries) == B:  # dA / dD sum over the whole batch: only comparable when every entry went through the oracle
        for k, v, r in (("dA", dA, dA_ref), ("dD", dD, dD_ref)):
            e = relerr(f64(v), r)
            if not (e <= tol):
                bad[k] = e
    assert not bad, f"rel-err above {tol} at B={B} L={L} ED={ED} {dtype} random_A={random_A} flags={flags:#x}: {bad}"


def test_bench_shape_fp32_all_entries():
    """bench.py's timed configuration (B=16, L=6400, d_inner=512, fp32, S4D-real A): every output of every batch entry,
    and dA / dD summed over the batch."""
    _check(16, 6400, 512, seed=1234)


def test_bench_shape_fp32_general_A():
    """the same shape with a trained (non-geometric) A: the 16-exponential path of both kernels."""
    _check(16, 6400, 512, random_A=True, entries=[0, 7, 15], seed=5)


def test_config0_shape():
    """BASELINE configs[0]: B=2, L=6400, d_inner=256, fp32 (the L-split schedule on a small grid)."""
    _check(2, 6400, 256, seed=11)
    _check(2, 6400, 256, random_A=True, seed=12)


@pytest.mark.parametrize("shape", [(16, 25600, 256), (16, 1600, 1024), (16, 400, 1024), (2, 25600, 512)])
def test_sweep_shapes_fp32(shape):
    """corner shapes of the configs[2] sweep (L = 400 ... 25600, d_inner = 256 ... 1024): three batch entries each."""
    B, L, ED = shape
    _check(B, L, ED, entries=[0, B // 2, B - 1], seed=L + ED)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("random_A", [False, True])
def test_bench_shape_16bit(dtype, random_A):
    """16-bit I/O at the bench shape: 6400-step carries in fp32 state against the fp64 oracle on the rounded inputs."""
    _check(16, 6400, 512, dtype=dtype, random_A=random_A, entries=[0, 9], seed=21)


@pytest.mark.parametrize("nseg", [1, 2, 4, 7])
def test_bench_shape_forced_segment_counts(nseg):
    """force the number of L segments (1 = one CTA walks all 6400 steps of its channel tile: 50 to 100 super-tiles through
    the mbarrier ring, the carry double buffer and the tensor-memory history slots) at a wide grid: values, both passes."""
    _check(8, 6400, 512, flags=nseg << 8, entries=[0, 7], seed=31 + nseg)


def test_no_gate_at_bench_shape():
    _check(4, 6400, 512, gate=False, entries=[1, 3], seed=41)
