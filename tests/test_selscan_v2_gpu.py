"""GPU parity of the second-generation scan kernels (selscan_fwd2.cu / selscan_bwd2.cu / selscan_bwd3.cu: persistent grid
over (32-channel chain, L segment) items, lanes split the states in halves or quarters, chained segments) forced with
MMI_FLAG_CFG = 8 (8-warp backward) and 11 (16-warp backward) on shapes that would otherwise take the first generation:
values against the fp64 C oracle.
Tolerances (north_star): 1e-4 fp32 I/O, 2e-2 bf16 I/O (5e-3 fp16), max|a-b| / max|b|."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from tests.util import relerr, scan_inputs

pytestmark = pytest.mark.gpu
V2 = 8 << 4


@pytest.fixture(autouse=True, params=[8, 11], ids=["w8", "w16"])
def _generation(request):
    """every test of this module runs once per backward kernel (the flag also selects the forward: 8 = second generation,
    11 = default dispatch)"""
    globals()["V2"] = request.param << 4
    yield
    globals()["V2"] = 8 << 4

TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2, torch.float16: 5e-3}


def _t(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda().to(dtype)


def _run(inp, dtype=torch.float32, gate=True, flags=None, softplus=False):
    from mmidet_b200 import _lib, ops
    flags = V2 if flags is None else flags
    if softplus:
        flags |= _lib.FLAG_DELTA_SOFTPLUS
    a = {k: _t(inp[k], torch.float32 if k in ("A", "D") else dtype) for k in ("x", "delta", "z", "A", "Bm", "Cm", "D", "dout")}
    out, _, chk, saved = ops.selscan_fwd_raw(a["x"], a["delta"], a["A"], a["Bm"], a["Cm"], a["D"], z=a["z"] if gate else None,
                                             want_chk=True, flags=flags)
    dx, dd, dz, dA, dB, dC, dD = ops.selscan_bwd_raw(saved, chk, a["dout"], flags=flags)
    torch.cuda.synchronize()
    res = dict(out=out, dx=dx, ddelta=dd, dA=dA, dB=dB, dC=dC, dD=dD)
    if gate:
        res["dz"] = dz
    return {k: v.float().cpu().numpy() for k, v in res.items()}


def _oracle(inp, gate=True):
    z = inp["z"] if gate else None
    ref = {"out": O.selective_scan_fwd(inp["x"], inp["delta"], inp["A"], inp["Bm"], inp["Cm"], inp["D"], z=z, dtype=np.float64)}
    ref.update(O.selective_scan_bwd(inp["x"], inp["delta"], inp["A"], inp["Bm"], inp["Cm"], inp["D"], inp["dout"], z=z,
                                    dtype=np.float64))
    return ref


def _compare(res, ref, tol):
    bad = {k: relerr(v, ref[k]) for k, v in res.items() if ref.get(k) is not None and not (relerr(v, ref[k]) <= tol)}
    assert not bad, f"rel-err above {tol}: {bad}"


@pytest.mark.parametrize("random_A", [False, True])
@pytest.mark.parametrize("shape", [(2, 96, 64), (1, 16, 8), (3, 37, 24), (2, 257, 40), (1, 1, 16), (2, 15, 72), (1, 300, 104),
                                   (2, 128, 32), (1, 129, 96), (3, 1000, 136), (1, 2049, 8)])
def test_v2_fp32_vs_oracle(shape, random_A):
    """ragged L (not a multiple of the 16-step chunk nor of the 128-step super-tile), ED not a multiple of the 32-channel
    chain, several super-tiles per chain, both A paths."""
    B, L, ED = shape
    inp = scan_inputs(B, L, ED, seed=L + ED, random_A=random_A)
    _compare(_run(inp), _oracle(inp), 1e-4)


@pytest.mark.parametrize("nseg", [1, 2, 3, 5, 16])
@pytest.mark.parametrize("random_A", [False, True])
def test_v2_chained_segments(nseg, random_A):
    """L cut into chained segments (carry handed through global memory in ticket order), more segments requested than
    super-tiles exist, more items than SMs (B * ED / 32 * nseg = 24 * nseg ... the persistent loop takes several items)."""
    B, L, ED = 3, 1500, 256
    inp = scan_inputs(B, L, ED, seed=nseg, random_A=random_A)
    _compare(_run(inp, flags=V2 | (nseg << 8)), _oracle(inp), 1e-4)


def test_v2_more_items_than_sms():
    """400 chains x 2 segments on 148 SMs: every CTA processes several items back to back (stage / parity / carry bookkeeping
    across item boundaries), dA / dD partials of every (batch, segment)."""
    B, L, ED = 50, 300, 256
    inp = scan_inputs(B, L, ED, seed=77)
    res = _run(inp, flags=V2 | (2 << 8))
    ref = _oracle(inp)
    _compare(res, ref, 1e-4)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("random_A", [False, True])
def test_v2_half_io(dtype, random_A):
    inp = scan_inputs(2, 700, 72, seed=11, random_A=random_A)
    rnd = {k: (torch.from_numpy(v).to(dtype).float().numpy() if k not in ("A", "D") else v) for k, v in inp.items()}
    _compare(_run(rnd, dtype=dtype, flags=V2 | (2 << 8)), _oracle(rnd), TOL[dtype])


@pytest.mark.parametrize("random_A", [False, True])
def test_v2_no_gate(random_A):
    inp = scan_inputs(2, 333, 40, seed=3, random_A=random_A)
    res = _run(inp, gate=False)
    assert "dz" not in res
    _compare(res, _oracle(inp, gate=False), 1e-4)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_v2_fused_softplus(dtype):
    B, L, ED = 2, 270, 40
    inp = scan_inputs(B, L, ED, seed=21, random_A=True)
    rng = np.random.default_rng(5)
    pre = (rng.standard_normal((B, L, ED)) * 2.0 - 3.0).astype(np.float32)
    pre[0, 3, :4] = [25.0, 19.5, -30.0, -12.0]
    cast = lambda a: torch.from_numpy(a).to(dtype).float().numpy()
    pre_r = cast(pre)
    sp = np.where(pre_r > 20, pre_r, np.log1p(np.exp(np.minimum(pre_r, 20)))).astype(np.float64)
    rnd = {k: (cast(v) if k not in ("A", "D") else v) for k, v in inp.items()}
    ref_in = dict(rnd)
    ref_in["delta"] = cast(sp.astype(np.float32)).astype(np.float64) if dtype != torch.float32 else sp
    ref = _oracle(ref_in)
    ref["ddelta"] = ref["ddelta"] * (1.0 / (1.0 + np.exp(-pre_r.astype(np.float64))))
    run_in = dict(rnd)
    run_in["delta"] = pre_r
    _compare(_run(run_in, dtype=dtype, softplus=True), ref, TOL[dtype])


def test_v2_large_delta_and_per_channel_base():
    B, L, ED = 2, 333, 72
    inp = scan_inputs(B, L, ED, seed=4)
    rng = np.random.default_rng(8)
    base = -(rng.random(ED).astype(np.float32) * 3.0 + 0.05)
    inp["A"] = (base[:, None] * np.arange(1, 17, dtype=np.float32)[None, :]).astype(np.float32)
    inp["delta"] = np.log1p(np.exp(rng.standard_normal((B, L, ED)) * 2.5)).astype(np.float32)
    ref = _oracle(inp)
    _compare(_run(inp), ref, 1e-4)
    _compare(_run(inp, flags=V2 | 1), ref, 1e-4)  # general 16-exponential path on the same data


def test_v2_is_bit_reproducible():
    """no atomics anywhere: two runs give identical bits whatever the SM-to-item assignment was."""
    inp = scan_inputs(4, 900, 200, seed=5, random_A=True)
    a, b = _run(inp, flags=V2 | (3 << 8)), _run(inp, flags=V2 | (3 << 8))
    for k in a:
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 77, 72), (1, 130, 8), (3, 200, 136), (1, 1000, 64)])
def test_v2_kernels_stay_inside_their_outputs(shape, dtype):
    """outputs and workspaces carved out of sentinel arenas (stands in for a memcheck tool)."""
    from mmidet_b200 import _lib, ops
    lib = _lib.load()
    P, DT, ST = ops._ptr, ops._DT, ops._stream
    B, L, ED = shape
    N, GUARD = 16, 4096
    arenas = []

    def carve(n, dt):
        buf = torch.full((n + 2 * GUARD,), 7.0 if dt != torch.uint8 else 7, device="cuda", dtype=dt)
        arenas.append((buf, n))
        return buf[GUARD:GUARD + n]

    torch.manual_seed(L)
    x, z, dout = (torch.randn(B, L, ED, device="cuda").to(dtype) for _ in range(3))
    delta = torch.nn.functional.softplus(torch.randn(B, L, ED, device="cuda") - 3).to(dtype)
    Bm, Cm = torch.randn(B, L, N, device="cuda").to(dtype), torch.randn(B, L, N, device="cuda").to(dtype)
    A = -torch.arange(1, N + 1, device="cuda", dtype=torch.float32).repeat(ED, 1).contiguous()
    D = torch.ones(ED, device="cuda")
    chunk = lib.mmi_selscan_chunk()
    nchk = (L + chunk - 1) // chunk
    fl = V2 | (2 << 8)
    out, hT, chk = carve(B * L * ED, dtype), carve(B * ED * N, torch.float32), carve(B * nchk * ED * N, torch.float32)
    wsf = carve(max(int(lib.mmi_selscan_fwd_ws_bytes(B, L, ED, N)), 16), torch.uint8)
    _lib.check(lib.mmi_selscan_fwd(P(x), P(delta), P(z), P(A), P(Bm), P(Cm), P(D), None, P(out), P(hT), P(chk), P(wsf), B, L, ED, N,
                                   ED, ED, ED, ED, chunk, DT[dtype], fl, ST(x)), "mmi_selscan_fwd")
    dx, dd, dz = carve(B * L * ED, dtype), carve(B * L * ED, dtype), carve(B * L * ED, dtype)
    dA, dD = carve(ED * N, torch.float32), carve(ED, torch.float32)
    dB, dC = carve(B * L * N, dtype), carve(B * L * N, dtype)
    wsb = carve(max(int(lib.mmi_selscan_bwd_ws_bytes(B, L, ED, N)), 16), torch.uint8)
    _lib.check(lib.mmi_selscan_bwd(P(x), P(delta), P(z), P(A), P(Bm), P(Cm), P(D), P(dout), P(chk), P(dx), P(dd), P(dz), P(dA), P(dB),
                                   P(dC), P(dD), P(wsb), B, L, ED, N, ED, ED, ED, ED, chunk, DT[dtype], fl, ST(x)), "mmi_selscan_bwd")
    torch.cuda.synchronize()
    for i, (buf, n) in enumerate(arenas):
        assert bool((buf[:GUARD] == 7).all()) and bool((buf[GUARD + n:] == 7).all()), f"guard band of arena {i} overwritten"
    for t in (out, dx, dd, dz, dA, dD, dB, dC):
        assert bool(torch.isfinite(t.float()).all())


@pytest.mark.parametrize("variant", [8, 10])  # 8 chunk-warps / one CTA per SM; 4 chunk-warps / two CTAs per SM
@pytest.mark.parametrize("nseg", [0, 1, 3])
@pytest.mark.parametrize("random_A", [False, True])
def test_v2_forward_matches_first_generation(nseg, random_A, variant):
    """outputs, final state hT and checkpoints of the second-generation forward == the first generation's, with a non-zero
    h0, a ragged tail and chained segments; and h0 -> hT chaining over a cut equals one call."""
    from mmidet_b200 import ops
    B, L, ED = 3, 1100, 136
    inp = scan_inputs(B, L, ED, seed=nseg + 40, random_A=random_A)
    a = {k: _t(v) for k, v in inp.items()}
    h0 = torch.randn(B, ED, 16, device="cuda")
    run = lambda fl, **kw: ops.selscan_fwd_raw(a["x"], a["delta"], a["A"], a["Bm"], a["Cm"], a["D"], z=a["z"], want_state=True,
                                               want_chk=True, flags=fl, **kw)
    o1, hT1, chk1, _ = run(9 << 4, h0=h0)
    V2v = variant << 4
    o2, hT2, chk2, _ = run(V2v | (nseg << 8), h0=h0)
    for u, v in ((o1, o2), (hT1, hT2), (chk1, chk2)):
        assert relerr(v.cpu().numpy(), u.cpu().numpy()) <= 2e-5
    ref_out, ref_h = O.selective_scan_fwd(inp["x"], inp["delta"], inp["A"], inp["Bm"], inp["Cm"], inp["D"], z=inp["z"],
                                          h0=h0.cpu().numpy(), dtype=np.float64, return_state=True)
    assert relerr(o2.cpu().numpy(), ref_out) <= 1e-4 and relerr(hT2.cpu().numpy(), ref_h) <= 1e-4
    cut = 517
    sl = lambda t, s: t[:, s].contiguous()
    first = ops.selscan_fwd_raw(sl(a["x"], slice(0, cut)), sl(a["delta"], slice(0, cut)), a["A"], sl(a["Bm"], slice(0, cut)),
                                sl(a["Cm"], slice(0, cut)), a["D"], z=sl(a["z"], slice(0, cut)), h0=h0, want_state=True, flags=V2v)
    second = ops.selscan_fwd_raw(sl(a["x"], slice(cut, L)), sl(a["delta"], slice(cut, L)), a["A"], sl(a["Bm"], slice(cut, L)),
                                 sl(a["Cm"], slice(cut, L)), a["D"], z=sl(a["z"], slice(cut, L)), h0=first[1], want_state=True,
                                 flags=V2v)
    assert relerr(torch.cat([first[0], second[0]], 1).cpu().numpy(), o2.cpu().numpy()) <= 1e-5
    assert relerr(second[1].cpu().numpy(), hT2.cpu().numpy()) <= 1e-5


def test_v2_strided_views():
    """x / z passed as chunk() views of one GEMM output (row pitch 2 * ED), as MambaBlock.forward produces them."""
    from mmidet_b200 import ops
    B, L, ED = 2, 300, 64
    inp = scan_inputs(B, L, ED, seed=5, random_A=True)
    xz = torch.cat([_t(inp["x"]), _t(inp["z"])], dim=-1)
    xv, zv = xz.chunk(2, dim=-1)
    out = ops.selscan_fwd_raw(xv, _t(inp["delta"]), _t(inp["A"]), _t(inp["Bm"]), _t(inp["Cm"]), _t(inp["D"]), z=zv, flags=V2)[0]
    assert relerr(out.cpu().numpy(), _oracle(inp)["out"]) <= 1e-4
