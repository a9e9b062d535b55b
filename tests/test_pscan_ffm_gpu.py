"""GPU parity for the pscan parity API and the FFM Fourier step (through the C ABI) vs goldens + oracle."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from tests.util import relerr

pytestmark = pytest.mark.gpu


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("tag", ["pow2", "ragged", "tiny"])
def test_pscan_vs_reference_golden(golden, tag):
    from mmidet_b200.pscan import pscan
    g = golden(f"pscan_{tag}")
    A, X = _t(g["A"]).requires_grad_(True), _t(g["X"]).requires_grad_(True)
    A0, X0 = A.detach().clone(), X.detach().clone()
    H = pscan(A, X)
    gA, gX = torch.autograd.grad(H, (A, X), _t(g["gH"]))
    assert relerr(H.detach().cpu().numpy(), g["H"]) <= 1e-5
    assert relerr(gA.cpu().numpy(), g["gA"]) <= 1e-5
    assert relerr(gX.cpu().numpy(), g["gX"]) <= 1e-5
    assert torch.equal(A.detach(), A0) and torch.equal(X.detach(), X0)  # inputs untouched (pscan.py:167-170)


@pytest.mark.parametrize("shape", [(2, 777, 32, 16), (1, 6400, 64, 16), (3, 65, 5, 3), (2, 1, 8, 16)])
def test_pscan_vs_oracle_multi_segment(shape):
    """non-pow2 L long enough to be cut into several segments; odd D*N."""
    from mmidet_b200.pscan import pscan
    rng = np.random.default_rng(1)
    A = (rng.random(shape) * 0.5 + 0.5).astype(np.float32)
    X = rng.standard_normal(shape).astype(np.float32)
    gH = rng.standard_normal(shape).astype(np.float32)
    At, Xt = _t(A).requires_grad_(True), _t(X).requires_grad_(True)
    H = pscan(At, Xt)
    gA, gX = torch.autograd.grad(H, (At, Xt), _t(gH))
    H64 = O.pscan_seq_fwd(A.astype(np.float64), X.astype(np.float64))
    gA64, gX64 = O.pscan_seq_bwd(A.astype(np.float64), H64, gH.astype(np.float64))
    assert relerr(H.detach().cpu().numpy(), H64) <= 1e-4
    assert relerr(gA.cpu().numpy(), gA64) <= 1e-4
    assert relerr(gX.cpu().numpy(), gX64) <= 1e-4


@pytest.mark.parametrize("tag", ["8x8", "16x16", "20x20", "7x7", "8x12"])
def test_ffm_vs_reference_golden(golden, tag):
    """extract_frequency2 incl. the negative-slice quirk; fp16 outputs may differ by one fp16 ulp."""
    from mmidet_b200.ffm import extract_frequency2, kept_range
    g = golden(f"ffm_{tag}")
    low, high, prod = extract_frequency2(_t(g["img"]), with_product=True)
    assert low.dtype == torch.float16 and high.dtype == torch.float16 and low.shape == g["img"].shape
    scale = float(np.abs(g["img"]).max())
    assert float(np.abs(low.float().cpu().numpy() - g["low"].astype(np.float32)).max()) <= 2e-3 * scale
    assert float(np.abs(high.float().cpu().numpy() - g["high"].astype(np.float32)).max()) <= 2e-3 * scale
    assert relerr(prod.cpu().numpy(), g["high"].astype(np.float32) * g["img"]) <= 4e-3
    # host-side kept range == the oracle's mask restatement
    H, W = g["img"].shape[-2:]
    r0, r1, c0, c1 = kept_range(H, W)
    _, keep_low = O.ffm_masks(H, W)
    m = np.zeros((H, W), bool)
    m[r0:r1, c0:c1] = True
    assert np.array_equal(m, keep_low)


def test_ffm_batch_and_dtypes():
    from mmidet_b200.ffm import extract_frequency2
    rng = np.random.default_rng(4)
    img = rng.standard_normal((4, 128, 8, 8)).astype(np.float32)
    lo_ref, hi_ref = O.extract_frequency2(img)
    for dt in (torch.float32, torch.float16, torch.bfloat16):
        x = _t(img).to(dt)
        lo, hi = extract_frequency2(x)
        l2, h2 = O.extract_frequency2(x.float().cpu().numpy())
        assert float((lo.float().cpu() - torch.from_numpy(l2.astype(np.float32))).abs().max()) <= 4e-3
        assert float((hi.float().cpu() - torch.from_numpy(h2.astype(np.float32))).abs().max()) <= 8e-3
    # size-independent properties: low + high reconstructs the input (complementary masks), and the split is linear
    lo, hi = extract_frequency2(_t(img))
    assert float((lo.float() + hi.float() - _t(img)).abs().max()) <= 4e-3
    lo2, hi2 = extract_frequency2(_t(2.0 * img))
    assert float((lo2.float() - 2 * lo.float()).abs().max()) <= 4e-3
    assert float((hi2.float() - 2 * hi.float()).abs().max()) <= 8e-3


def test_separation_loss(golden):
    from mmidet_b200.ffm import separation_loss
    g = golden("seploss")
    v = float(separation_loss(_t(g["M"])))
    assert abs(v - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    M = np.random.default_rng(0).random((288, 64)).astype(np.float32)  # l = 18 * 16 (B=16), SURVEY 8a
    assert abs(float(separation_loss(_t(M))) - O.separation_loss(M)) <= 1e-4 * O.separation_loss(M)
