"""CPU: pin the oracle (oracle/) against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  Tolerances are stated per test; fp32 reference vs fp32 restatement
differs only by summation order / libm exp."""
import numpy as np
import pytest

from oracle import oracle as O


def relerr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


@pytest.mark.parametrize("tag", ["pow2", "ragged", "tiny", "f64"])
def test_pscan_blelloch_matches_reference(golden, tag):
    """numpy restatement of the Blelloch sweeps (pscan.py:37-224) == reference, bit-for-bit op order."""
    g = golden(f"pscan_{tag}")
    H, Hpad = O.pscan_forward(g["A"], g["X"])
    tol = 1e-12 if tag == "f64" else 2e-6
    assert relerr(H, g["H"]) <= tol
    gA, gX = O.pscan_backward(g["A"], Hpad, g["gH"])
    assert relerr(gA, g["gA"]) <= tol
    assert relerr(gX, g["gX"]) <= tol


@pytest.mark.parametrize("tag", ["pow2", "ragged", "tiny", "f64"])
def test_pscan_sequential_matches_reference(golden, tag):
    """sequential C recurrence == reference pscan fwd/bwd (1.3e-7 fp32 / 2e-16 fp64)."""
    g = golden(f"pscan_{tag}")
    tol = 1e-12 if tag == "f64" else 5e-6
    H = O.pscan_seq_fwd(g["A"], g["X"])
    assert relerr(H, g["H"]) <= tol
    gA, gX = O.pscan_seq_bwd(g["A"], H, g["gH"])
    assert relerr(gA, g["gA"]) <= tol
    assert relerr(gX, g["gX"]) <= tol


@pytest.mark.parametrize("tag", ["init", "randA", "short"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_selscan_fwd_matches_reference(golden, tag, dtype):
    g = golden(f"selscan_{tag}")
    y = O.selective_scan_fwd(g["x"], g["delta"], g["A"], g["Bm"], g["Cm"], g["D"], dtype=dtype)
    assert relerr(y, g["y_pscan"]) <= 5e-6
    assert relerr(y, g["y_seq"]) <= 5e-6
    out = O.selective_scan_fwd(g["x"], g["delta"], g["A"], g["Bm"], g["Cm"], g["D"], z=g["z"], dtype=dtype)
    assert relerr(out, g["out"]) <= 5e-6
    y2 = O.selective_scan_pscan(g["x"], g["delta"], g["A"], g["Bm"], g["Cm"], g["D"])
    assert relerr(y2, g["y_pscan"]) <= 5e-6


@pytest.mark.parametrize("tag", ["init", "randA", "short"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_selscan_bwd_matches_reference_autograd(golden, tag, dtype):
    g = golden(f"selscan_{tag}")
    r = O.selective_scan_bwd(g["x"], g["delta"], g["A"], g["Bm"], g["Cm"], g["D"], g["dout"], z=g["z"], dtype=dtype)
    for k in ("dx", "ddelta", "dz", "dA", "dB", "dC", "dD"):
        assert relerr(r[k], g[k]) <= 2e-5, k


def test_selscan_state_passing():
    """splitting L with (h0 -> hT) state passing reproduces the unsplit scan (used by the L-split path)."""
    rng = np.random.default_rng(0)
    B, L, ED, N = 2, 50, 8, 16
    x, z = rng.standard_normal((2, B, L, ED))
    delta = np.log1p(np.exp(rng.standard_normal((B, L, ED)) - 3))
    A = -np.exp(rng.standard_normal((ED, N)) * 0.5)
    Bm, Cm = rng.standard_normal((2, B, L, N))
    D = rng.standard_normal(ED)
    full = O.selective_scan_fwd(x, delta, A, Bm, Cm, D, z=z, dtype=np.float64)
    a, h = O.selective_scan_fwd(x[:, :20], delta[:, :20], A, Bm[:, :20], Cm[:, :20], D, z=z[:, :20], dtype=np.float64,
                                return_state=True)
    b = O.selective_scan_fwd(x[:, 20:], delta[:, 20:], A, Bm[:, 20:], Cm[:, 20:], D, z=z[:, 20:], h0=h,
                             dtype=np.float64)
    assert relerr(np.concatenate([a, b], 1), full) <= 1e-13


def test_mamba_block_matches_reference(golden):
    g = golden("mamba_block")
    p = {k[len("sd.mixer."):]: v for k, v in g.items() if k.startswith("sd.mixer.")}
    y = O.mamba_block_forward(g["x"], p)
    assert relerr(y, g["y_mixer"]) <= 1e-5
    yr = O.mamba_block_forward(O.rmsnorm(g["x"], g["sd.norm.weight"]), p) + g["x"]
    assert relerr(yr, g["y_res"]) <= 1e-5


def test_step_matches_forward_and_reference(golden):
    """MambaBlock.step / ssm_step (mamba.py:289-353): the recurrent form over 8 tokens == forward() on the same prefix
    (reference outputs), and the oracle's block restatement reproduces both."""
    g = golden("mamba_step")
    assert relerr(g["y_step"], g["y_fwd"]) <= 1e-5
    p = {k[len("sd.mixer."):]: v for k, v in g.items() if k.startswith("sd.mixer.")}
    y = O.mamba_block_forward(O.rmsnorm(g["x"], g["sd.norm.weight"]), p) + g["x"]
    assert relerr(y, g["y_step"]) <= 1e-5


@pytest.mark.parametrize("tag", ["8x8", "80x80"])
def test_extract_frequency_matches_reference(golden, tag):
    """extract_frequency (common.py:72-93; no caller in the reference): fixed threshold 30, complex -> real -> fp16."""
    g = golden(f"ffm_{tag}")
    lo, hi = O.extract_frequency(g["img"].astype(np.float32))
    scale = max(float(np.max(np.abs(g["ef_high"]))), 1e-6)
    assert np.max(np.abs(np.asarray(lo, np.float32) - g["ef_low"])) <= 2e-3 * scale
    assert np.max(np.abs(np.asarray(hi, np.float32) - g["ef_high"])) <= 2e-3 * scale


@pytest.mark.parametrize("tag", ["8x8", "16x16", "20x20", "7x7", "8x12", "80x80", "160x160", "96x72"])
def test_ffm_matches_reference(golden, tag):
    """extract_frequency2 incl. the negative-slice wrap (common.py:44-56) and the fp16 real cast (:66-67)."""
    g = golden(f"ffm_{tag}")
    g["img"] = g["img"].astype(np.float32)
    low, high = O.extract_frequency2(g["img"])
    assert low.dtype == np.float16 and high.dtype == np.float16
    # fp16 outputs: allow 1 fp16 ulp of the largest magnitude (pocketfft vs numpy fft rounding)
    assert relerr(low, g["low"]) <= 2e-3
    assert relerr(high, g["high"]) <= 2e-3
    if "fs_re" in g:
        fs = O.fourier_transform(g["img"])
        assert relerr(fs.real, g["fs_re"]) <= 1e-5 and relerr(fs.imag, g["fs_im"]) <= 1e-5
    # mask restatement == slice restatement
    kh, kl = O.ffm_masks(*g["img"].shape[-2:])
    fsh = np.fft.fftshift(np.fft.fftn(g["img"], axes=(-2, -1)), axes=(-2, -1))
    hi2 = np.fft.ifftn(np.fft.ifftshift(fsh * kh, axes=(-2, -1)), axes=(-2, -1)).real
    lo2 = np.fft.ifftn(np.fft.ifftshift(fsh * kl, axes=(-2, -1)), axes=(-2, -1)).real
    assert relerr(hi2, g["high"].astype(np.float32)) <= 2e-3
    assert relerr(lo2, g["low"].astype(np.float32)) <= 2e-3 or np.max(np.abs(g["low"])) < 1e-3


def test_separation_loss_matches_reference(golden):
    g = golden("seploss")
    assert abs(O.separation_loss(g["M"]) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))


@pytest.mark.parametrize("tag", ["b2", "b9", "gpt1"])
def test_ffm_pattern_matches_reference(golden, tag):
    """pattern path of GPT1_fourier.forward (common.py:434-516) / GPT1.forward (:218-262) vs tensors captured inside the
    unmodified reference."""
    g = golden(f"pattern_{tag}")
    C = g["pool_vis"].shape[1]
    tok, loss = O.ffm_pattern(g["pool_vis"], g["pool_ir"], g["conv1_w"].reshape(8, C), g["conv2_w"].reshape(C, 8),
                              high=tag != "gpt1")
    assert relerr(tok + g["pos_emb"], g["drop_in"]) <= 1e-5
    assert abs(loss - float(g["loss"][0])) <= 1e-5 * abs(float(g["loss"][0]))


@pytest.mark.parametrize("shape,anchors", [((2, 3, 12, 10), (8, 8)), ((1, 2, 37, 53), (8, 8)), ((1, 1, 160, 160), (8, 8)),
                                           ((2, 2, 8, 8), (8, 8)), ((1, 3, 9, 11), (1, 1)), ((1, 1, 33, 20), (4, 6))])
def test_resample_oracle_matches_torch(shape, anchors):
    """the pooling windows and bilinear taps the kernels implement == the torch ops the reference calls
    (nn.AdaptiveAvgPool2d, common.py:324-325; F.interpolate bilinear, common.py:540-543)."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(7)
    x = rng.standard_normal(shape)
    s = rng.standard_normal(shape[:2] + anchors)
    assert relerr(O.adaptive_avg_pool2d(x, anchors), F.adaptive_avg_pool2d(torch.from_numpy(x), anchors).numpy()) <= 1e-12
    want = F.interpolate(torch.from_numpy(s), size=list(shape[2:]), mode="bilinear").numpy()
    assert relerr(O.upsample_bilinear(s, shape[2:]), want) <= 1e-12


def test_causal_conv_oracle_matches_torch_conv1d():
    """the oracle's conv restatement vs the exact torch ops of models/mamba.py:176-180 (Conv1d padding=K-1, [:L], silu)."""
    import torch
    import torch.nn.functional as F
    torch.manual_seed(0)
    for K in (4, 2):
        B, L, ED = 2, 19, 8
        conv = torch.nn.Conv1d(ED, ED, kernel_size=K, groups=ED, padding=K - 1).double()
        x = torch.randn(B, L, ED, dtype=torch.float64, requires_grad=True)
        y = F.silu(conv(x.transpose(1, 2))[:, :, :L].transpose(1, 2))
        g = torch.randn_like(y)
        gx, gw, gb = torch.autograd.grad(y, [x, conv.weight, conv.bias], g)
        w = conv.weight.detach().numpy()[:, 0, :]
        b = conv.bias.detach().numpy()
        assert np.allclose(O.causal_conv1d_silu(x.detach().numpy(), w, b), y.detach().numpy(), atol=1e-12)
        dx, dw, db = O.causal_conv1d_silu_bwd(x.detach().numpy(), w, b, g.numpy())
        assert np.allclose(dx, gx.numpy(), atol=1e-12) and np.allclose(dw, gw.numpy()[:, 0, :], atol=1e-12)
        assert np.allclose(db, gb.numpy(), atol=1e-12)
