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


@pytest.mark.parametrize("shape", [(2, 777, 32, 16), (1, 6400, 64, 16), (3, 65, 5, 3), (2, 1, 8, 16), (1, 31, 9, 16), (2, 32, 8, 16),
                                   (1, 33, 3, 16), (2, 2000, 9, 16), (64, 200, 4, 16)])
def test_pscan_vs_oracle_multi_segment(shape):
    """non-pow2 L cut into many 32-step segments (look-back chains of up to 200 records), L around the segment length, D*N
    below / not a multiple of the 128 columns of a CTA, more (batch, column block) chains than resident CTAs."""
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


def test_pscan_one_long_chain():
    """a single (batch, column block) chain of 2000 segments: every resident CTA belongs to the same chain, so the look-back
    regularly runs into its depth limit and has to wait for an inclusive state (the path the wide shapes rarely take)."""
    from mmidet_b200.pscan import pscan
    rng = np.random.default_rng(12)
    shape = (1, 64000, 8, 16)
    A = (rng.random(shape) * 0.2 + 0.8).astype(np.float32)
    X = rng.standard_normal(shape).astype(np.float32)
    gH = rng.standard_normal(shape).astype(np.float32)
    At, Xt = _t(A).requires_grad_(True), _t(X).requires_grad_(True)
    H = pscan(At, Xt)
    gA, gX = torch.autograd.grad(H, (At, Xt), _t(gH))
    H64 = O.pscan_seq_fwd(A.astype(np.float64), X.astype(np.float64))
    gA64, gX64 = O.pscan_seq_bwd(A.astype(np.float64), H64, gH.astype(np.float64))
    assert relerr(H.detach().cpu().numpy(), H64) <= 1e-4
    assert relerr(gA.cpu().numpy(), gA64) <= 1e-4 and relerr(gX.cpu().numpy(), gX64) <= 1e-4
    H2 = pscan(At, Xt)
    assert torch.equal(H.detach(), H2.detach())


def test_pscan_is_bit_reproducible():
    """the look-back chains the parked aggregates oldest first, so the value does not depend on how deep each thread had to
    look: repeated runs give identical bits, forward and backward."""
    from mmidet_b200.pscan import pscan
    rng = np.random.default_rng(4)
    shape = (2, 3000, 40, 16)
    A = _t((rng.random(shape) * 0.5 + 0.5).astype(np.float32)).requires_grad_(True)
    X = _t(rng.standard_normal(shape).astype(np.float32)).requires_grad_(True)
    gH = _t(rng.standard_normal(shape).astype(np.float32))
    runs = []
    for _ in range(4):
        H = pscan(A, X)
        gA, gX = torch.autograd.grad(H, (A, X), gH)
        runs.append((H.detach().clone(), gA.clone(), gX.clone()))
    for r in runs[1:]:
        for u, v in zip(runs[0], r):
            assert torch.equal(u, v)


def test_pscan_fp64_vs_reference_golden(golden):
    """models/pscan.py is dtype-generic: float64 inputs are scanned in float64 (fixture from the reference's own fp64 run)."""
    from mmidet_b200.pscan import pscan
    g = golden("pscan_f64")
    A, X = _t(g["A"]).requires_grad_(True), _t(g["X"]).requires_grad_(True)
    assert A.dtype == torch.float64
    H = pscan(A, X)
    gA, gX = torch.autograd.grad(H, (A, X), _t(g["gH"]))
    assert H.dtype == torch.float64 and gA.dtype == torch.float64
    assert relerr(H.detach().cpu().numpy(), g["H"]) <= 1e-12
    assert relerr(gA.cpu().numpy(), g["gA"]) <= 1e-12
    assert relerr(gX.cpu().numpy(), g["gX"]) <= 1e-12


def test_pscan_fp64_multi_segment():
    from mmidet_b200.pscan import pscan
    rng = np.random.default_rng(3)
    shape = (2, 1500, 12, 16)
    A = rng.random(shape) * 0.5 + 0.5
    X = rng.standard_normal(shape)
    gH = rng.standard_normal(shape)
    At, Xt = _t(A).requires_grad_(True), _t(X).requires_grad_(True)
    H = pscan(At, Xt)
    gA, gX = torch.autograd.grad(H, (At, Xt), _t(gH))
    H64 = O.pscan_seq_fwd(A, X)
    gA64, gX64 = O.pscan_seq_bwd(A, H64, gH)
    assert relerr(H.detach().cpu().numpy(), H64) <= 1e-12
    assert relerr(gA.cpu().numpy(), gA64) <= 1e-12 and relerr(gX.cpu().numpy(), gX64) <= 1e-12


def test_pscan_config0_shape():
    """BASELINE configs[0] as the materialised parity API: (B, L, D, N) = (2, 6400, 256, 16), 210 MB per tensor, ~70 segments per
    sequence (forward and reverse); values against the sequential fp64 oracle."""
    from mmidet_b200.pscan import pscan
    rng = np.random.default_rng(9)
    shape = (2, 6400, 256, 16)
    A = (rng.random(shape, dtype=np.float32) * 0.3 + 0.7)
    X = rng.standard_normal(shape, dtype=np.float32)
    gH = rng.standard_normal(shape, dtype=np.float32)
    At, Xt = _t(A).requires_grad_(True), _t(X).requires_grad_(True)
    H = pscan(At, Xt)
    gA, gX = torch.autograd.grad(H, (At, Xt), _t(gH))
    for b in range(2):  # one batch entry at a time keeps the fp64 oracle's footprint small
        A64, X64, g64 = A[b:b + 1].astype(np.float64), X[b:b + 1].astype(np.float64), gH[b:b + 1].astype(np.float64)
        H64 = O.pscan_seq_fwd(A64, X64)
        gA64, gX64 = O.pscan_seq_bwd(A64, H64, g64)
        assert relerr(H[b:b + 1].detach().cpu().numpy(), H64) <= 1e-4
        assert relerr(gA[b:b + 1].cpu().numpy(), gA64) <= 1e-4
        assert relerr(gX[b:b + 1].cpu().numpy(), gX64) <= 1e-4


@pytest.mark.parametrize("tag", ["80x80", "160x160", "96x72"])
def test_ffm_large_maps_vs_reference_golden(golden, tag):
    """extract_frequency2 above 64 x 64 (H in {8, 16, 20, 80, 160}): the kept-bin projection kernel
    against outputs of the unmodified reference; fp16 outputs, one fp16 ulp of the largest magnitude allowed."""
    from mmidet_b200.ffm import extract_frequency2
    g = golden(f"ffm_{tag}")
    img = _t(g["img"].astype(np.float32))
    low, high, prod = extract_frequency2(img, with_product=True)
    assert low.dtype == torch.float16 and high.dtype == torch.float16 and low.shape == img.shape
    assert relerr(low.float().cpu().numpy(), g["low"].astype(np.float32)) <= 2e-3
    assert relerr(high.float().cpu().numpy(), g["high"].astype(np.float32)) <= 2e-3
    assert relerr(prod.cpu().numpy(), high.float().cpu().numpy() * g["img"].astype(np.float32)) <= 1e-6
    lo16, hi16 = extract_frequency2(img.half())  # fp16 callers (detect_twostream.py:45 model.half())
    assert relerr(lo16.float().cpu().numpy(), g["low"].astype(np.float32)) <= 4e-3


@pytest.mark.parametrize("tag", ["8x8", "16x16", "20x20", "8x12"])
def test_fourier_transform_vs_reference_golden(golden, tag):
    """fourier_transform (models/common.py:25-32) on the device against the reference's spectrum."""
    from mmidet_b200.ffm import fourier_transform
    g = golden(f"ffm_{tag}")
    fs = fourier_transform(_t(g["img"]))
    assert relerr(fs.real.cpu().numpy(), g["fs_re"]) <= 1e-5 and relerr(fs.imag.cpu().numpy(), g["fs_im"]) <= 1e-5


@pytest.mark.parametrize("tag", ["8x8", "80x80"])
def test_extract_frequency_vs_reference_golden(golden, tag):
    """extract_frequency (models/common.py:72-93, no caller in the reference): fixed threshold 30, spectra cast to fp16."""
    from mmidet_b200.ffm import extract_frequency
    g = golden(f"ffm_{tag}")
    lo, hi = extract_frequency(_t(g["img"].astype(np.float32)))
    assert lo.dtype == torch.float16 and hi.dtype == torch.float16
    scale = max(float(np.max(np.abs(g["ef_high"]))), 1e-6)
    assert np.max(np.abs(lo.float().cpu().numpy() - g["ef_low"])) <= 2e-3 * scale
    assert np.max(np.abs(hi.float().cpu().numpy() - g["ef_high"])) <= 2e-3 * scale


def test_ffm_helpers_refuse_silent_zero_gradients():
    """the value-only drop-ins raise when asked to differentiate instead of returning no grad_fn."""
    from mmidet_b200.ffm import extract_frequency2, separation_loss
    img = torch.randn(1, 2, 8, 8, device="cuda", requires_grad=True)
    with pytest.raises(RuntimeError):
        extract_frequency2(img)
    with torch.no_grad():
        extract_frequency2(img)
    with pytest.raises(RuntimeError):
        separation_loss(torch.rand(6, 64, device="cuda", requires_grad=True))


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


# ---------------------------------------------------------------------------------------------------------
# FFM pattern path (GPT1_fourier.forward between pooling and transformer, models/common.py:434-516)
# ---------------------------------------------------------------------------------------------------------
def _pattern_torch(vis, ir, w1, w2):
    """plain torch restatement of the token path (differentiable, any dtype) used for gradient parity."""
    toks = []
    for fea in (vis, ir):
        M = torch.sigmoid(torch.einsum("jc,bchw->bjhw", w1, fea))
        toks.append((torch.einsum("cj,bjhw->bchw", w2, M) * fea).flatten(2))
    return torch.cat(toks, dim=2).transpose(1, 2)


@pytest.mark.parametrize("tag", ["b2", "b9", "gpt1"])
def test_pattern_tokens_vs_reference_golden(golden, tag):
    """tokens entering self.drop and pattenLoss captured inside the unmodified GPT1_fourier.forward / GPT1.forward."""
    from mmidet_b200.ffm import pattern_tokens
    g = golden(f"pattern_{tag}")
    tok, loss = pattern_tokens(_t(g["pool_vis"]), _t(g["pool_ir"]), _t(g["conv1_w"]), _t(g["conv2_w"]), high=tag != "gpt1")
    assert tok.shape == g["drop_in"].shape and loss.dim() == 0 and not loss.requires_grad
    assert relerr(tok.cpu().numpy() + g["pos_emb"], g["drop_in"]) <= 1e-5
    assert abs(float(loss) - float(g["loss"][0])) <= 1e-5 * abs(float(g["loss"][0]))


class _FourierStandIn(torch.nn.Module):
    """Attribute-for-attribute stand-in for GPT1_fourier(d_model, n_layer=0) (models/common.py:300-343): the
    reference class does not travel to the GPU box, fourier_forward only touches these members."""

    def __init__(self, g):
        super().__init__()
        C = g["conv1_w"].shape[1]
        self.n_embd, self.vert_anchors, self.horz_anchors = C, 8, 8
        self.pos_emb = torch.nn.Parameter(torch.from_numpy(g["pos_emb"]))
        self.trans_blocks = torch.nn.Sequential()
        self.ln_f = torch.nn.LayerNorm(C)
        self.drop = torch.nn.Dropout(0.1)
        self.avgpool = torch.nn.AdaptiveAvgPool2d((8, 8))
        self.conv1 = torch.nn.Conv2d(C, 8, kernel_size=1, bias=False)
        self.conv2 = torch.nn.Conv2d(8, C, kernel_size=1, bias=False)
        with torch.no_grad():
            self.ln_f.weight.copy_(torch.from_numpy(g["ln_w"]))
            self.ln_f.bias.copy_(torch.from_numpy(g["ln_b"]))
            self.conv1.weight.copy_(torch.from_numpy(g["conv1_w"]))
            self.conv2.weight.copy_(torch.from_numpy(g["conv2_w"]))


@pytest.mark.parametrize("tag", ["b2", "b9", "gpt1"])
def test_fourier_forward_vs_reference_golden(golden, tag):
    """the replacement bound onto GPT1_fourier.forward: both output maps, the loss and the gradients of a seeded
    functional of the outputs w.r.t. both inputs, conv1, conv2 and pos_emb -- all from the unmodified reference."""
    from mmidet_b200 import ffm
    fourier_forward = ffm.gpt1_forward if tag == "gpt1" else ffm.fourier_forward  # GPT1: the sibling without the Fourier branch
    g = golden(f"pattern_{tag}")
    m = _FourierStandIn(g).cuda().eval()
    vis, ir = _t(g["vis"]).requires_grad_(True), _t(g["ir"]).requires_grad_(True)
    ro, io, loss = fourier_forward(m, [vis, ir])
    assert m.pattenLoss is loss
    assert relerr(ro.detach().cpu().numpy(), g["rgb_out"]) <= 1e-4
    assert relerr(io.detach().cpu().numpy(), g["ir_out"]) <= 1e-4
    assert abs(float(loss) - float(g["loss"][0])) <= 1e-5 * abs(float(g["loss"][0]))
    ((ro * _t(g["g1"])).sum() + (io * _t(g["g2"])).sum()).backward()
    for got, key in ((vis.grad, "d_vis"), (ir.grad, "d_ir"), (m.conv1.weight.grad, "d_conv1"),
                     (m.conv2.weight.grad, "d_conv2"), (m.pos_emb.grad, "d_pos")):
        assert relerr(got.cpu().numpy(), g[key]) <= 1e-4, key


@pytest.mark.parametrize("B,C,hw", [(1, 8, (8, 8)), (3, 100, (4, 6)), (16, 256, (8, 8)), (2, 1024, (8, 16)), (17, 33, (5, 5))])
def test_pattern_tokens_vs_oracle_and_autograd(B, C, hw):
    """forward vs the numpy oracle, backward vs fp64 autograd of the same formula; ragged C, P and B > 8 (two
    batch entries feed the high-pass rows)."""
    from mmidet_b200.ffm import pattern_tokens
    rng = np.random.default_rng(B * 1000 + C)
    vis = rng.standard_normal((B, C, *hw)).astype(np.float32)
    ir = (rng.standard_normal((B, C, *hw)) * 0.5 + 0.3).astype(np.float32)
    w1 = (rng.standard_normal((8, C)) * 2.0 / np.sqrt(C)).astype(np.float32)
    w2 = rng.standard_normal((C, 8)).astype(np.float32)
    dtok = rng.standard_normal((B, 2 * hw[0] * hw[1], C)).astype(np.float32)
    tv, ti = _t(vis).requires_grad_(True), _t(ir).requires_grad_(True)
    t1, t2 = _t(w1.reshape(8, C, 1, 1)).requires_grad_(True), _t(w2.reshape(C, 8, 1, 1)).requires_grad_(True)
    tok, loss = pattern_tokens(tv, ti, t1, t2)
    otok, oloss = O.ffm_pattern(vis, ir, w1, w2)
    assert relerr(tok.detach().cpu().numpy(), otok) <= 1e-5
    assert abs(float(loss) - oloss) <= 1e-5 * abs(oloss)
    grads = torch.autograd.grad(tok, (tv, ti, t1, t2), _t(dtok))
    dv, di, d1, d2 = (torch.from_numpy(a).double().cuda().requires_grad_(True) for a in (vis, ir, w1, w2))
    ref = torch.autograd.grad(_pattern_torch(dv, di, d1, d2), (dv, di, d1, d2), _t(dtok).double())
    for got, want, name in zip(grads, ref, ("dvis", "dir", "dconv1", "dconv2")):
        assert relerr(got.cpu().numpy().reshape(want.shape), want.cpu().numpy()) <= 1e-4, name


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_pattern_tokens_half_inputs(dtype):
    """16-bit pooled maps (autocast): fp32 arithmetic inside, tokens and input gradients in the input dtype."""
    from mmidet_b200.ffm import pattern_tokens
    torch.manual_seed(3)
    B, C = 4, 128
    vis = torch.randn(B, C, 8, 8, device="cuda").to(dtype).requires_grad_(True)
    ir = torch.randn(B, C, 8, 8, device="cuda").to(dtype).requires_grad_(True)
    w1 = (torch.randn(8, C, 1, 1, device="cuda") * 0.2).requires_grad_(True)
    w2 = torch.randn(C, 8, 1, 1, device="cuda").requires_grad_(True)
    dtok = torch.randn(B, 128, C, device="cuda")
    tok, loss = pattern_tokens(vis, ir, w1, w2)
    assert tok.dtype == dtype
    grads = torch.autograd.grad(tok, (vis, ir, w1, w2), dtok.to(dtype))
    dv, di = vis.detach().double().requires_grad_(True), ir.detach().double().requires_grad_(True)
    d1, d2 = w1.detach().double().reshape(8, C).requires_grad_(True), w2.detach().double().reshape(C, 8).requires_grad_(True)
    rt = _pattern_torch(dv, di, d1, d2)
    ref = torch.autograd.grad(rt, (dv, di, d1, d2), dtok.to(dtype).double())
    assert relerr(tok.detach().float().cpu().numpy(), rt.detach().cpu().numpy()) <= 2e-2
    otok, oloss = O.ffm_pattern(vis.detach().float().cpu().numpy(), ir.detach().float().cpu().numpy(),
                                w1.detach().reshape(8, C).cpu().numpy(), w2.detach().reshape(C, 8).cpu().numpy())
    assert abs(float(loss) - oloss) <= 1e-4 * abs(oloss)
    for got, want in zip(grads, ref):
        assert relerr(got.float().cpu().numpy().reshape(want.shape), want.cpu().numpy()) <= 2e-2


def test_pattern_tokens_rejects_bad_input():
    from mmidet_b200.ffm import pattern_tokens
    with pytest.raises(RuntimeError):
        pattern_tokens(torch.zeros(1, 8, 8, 8), torch.zeros(1, 8, 8, 8), torch.zeros(8, 8, 1, 1), torch.zeros(8, 8, 1, 1))
    z = torch.zeros(1, 8, 12, 12, device="cuda")  # 144 pooled positions > 128
    with pytest.raises(RuntimeError, match="vert_anchors"):
        pattern_tokens(z, z, torch.zeros(8, 8, 1, 1, device="cuda"), torch.zeros(8, 8, 1, 1, device="cuda"))


# ---------------------------------------------------------------------------------------------------------
# resampling either side of the FFM token path (models/common.py:324-325, :396-397, :540-543)
# ---------------------------------------------------------------------------------------------------------
_RESAMPLE_SHAPES = [((2, 3, 160, 160), (8, 8)), ((1, 5, 37, 53), (8, 8)), ((2, 4, 20, 24), (4, 6)), ((1, 2, 8, 8), (8, 8)),
                    ((1, 3, 12, 10), (8, 8)), ((1, 1, 333, 64), (16, 16)), ((3, 2, 9, 11), (1, 1)),
                    # >= 2 images per SM and >= 32 KB per image: the shared-memory ring variant of the reduce kernel
                    ((4, 80, 96, 96), (8, 8)), ((2, 150, 128, 136), (8, 8)), ((5, 67, 100, 92), (5, 7))]


@pytest.mark.parametrize("shape,anchors", _RESAMPLE_SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_adaptive_avg_pool_matches_torch(shape, anchors, dtype):
    """forward and backward vs nn.AdaptiveAvgPool2d (the op the reference calls), incl. overlapping ragged windows."""
    import torch.nn.functional as F
    from mmidet_b200 import ops
    torch.manual_seed(1)
    x = torch.randn(*shape, device="cuda").to(dtype).requires_grad_(True)
    g = torch.randn(*shape[:2], *anchors, device="cuda").to(dtype)
    y = ops.adaptive_avg_pool(x, anchors)
    (dx,) = torch.autograd.grad(y, x, g)
    xr = x.detach().double().requires_grad_(True)
    yr = F.adaptive_avg_pool2d(xr, anchors)
    (dxr,) = torch.autograd.grad(yr, xr, g.double())
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert y.dtype == dtype and dx.shape == x.shape
    assert relerr(y.detach().float().cpu().numpy(), yr.detach().cpu().numpy()) <= tol
    assert relerr(dx.float().cpu().numpy(), dxr.cpu().numpy()) <= tol


@pytest.mark.parametrize("shape,anchors", _RESAMPLE_SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_upsample_bilinear_matches_torch(shape, anchors, dtype):
    """forward and backward vs F.interpolate(mode='bilinear') with the default align_corners=False."""
    import torch.nn.functional as F
    from mmidet_b200 import ops
    torch.manual_seed(2)
    x = torch.randn(*shape[:2], *anchors, device="cuda").to(dtype).requires_grad_(True)
    g = torch.randn(*shape, device="cuda").to(dtype)
    y = ops.upsample_bilinear(x, shape[2:])
    (dx,) = torch.autograd.grad(y, x, g)
    xr = x.detach().double().requires_grad_(True)
    yr = F.interpolate(xr, size=list(shape[2:]), mode="bilinear")
    (dxr,) = torch.autograd.grad(yr, xr, g.double())
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert y.shape == tuple(shape) and y.dtype == dtype
    assert relerr(y.detach().float().cpu().numpy(), yr.detach().cpu().numpy()) <= tol
    assert relerr(dx.float().cpu().numpy(), dxr.cpu().numpy()) <= tol


def test_resample_rejects_unsupported():
    from mmidet_b200 import ops
    with pytest.raises(RuntimeError):
        ops.adaptive_avg_pool(torch.zeros(1, 1, 8, 8), (8, 8))
    with pytest.raises(RuntimeError, match="at least as large"):
        ops.adaptive_avg_pool(torch.zeros(1, 1, 4, 4, device="cuda"), (8, 8))
    with pytest.raises(RuntimeError, match="anchor grid"):
        ops.upsample_bilinear(torch.zeros(1, 1, 32, 32, device="cuda"), (64, 64))


def test_fourier_forward_graph_replay(golden):
    """batch-1 inference of the FFM forward replayed as a CUDA graph (graphs.Graphed) == eager, for fresh inputs too."""
    from mmidet_b200.ffm import fourier_forward
    from mmidet_b200.graphs import Graphed
    g = golden("pattern_b2")
    m = _FourierStandIn(g).cuda().eval()
    C = m.n_embd
    torch.manual_seed(5)
    m.trans_blocks = torch.nn.Sequential(torch.nn.Linear(C, C), torch.nn.GELU(), torch.nn.Linear(C, C)).cuda()

    class Wrap(torch.nn.Module):
        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def forward(self, x):
            return fourier_forward(self.inner, x)

    fast = Graphed(Wrap(m))
    for seed in (0, 1, 2):
        torch.manual_seed(seed)
        vis, ir = torch.randn(1, C, 40, 40, device="cuda"), torch.randn(1, C, 40, 40, device="cuda")
        with torch.no_grad():
            want = fourier_forward(m, [vis, ir])
        got = fast([vis, ir])
        for a, b in zip(got, want):
            assert torch.allclose(a, b, rtol=0, atol=0), seed


def test_resample_fuzz_vs_oracle():
    """random map / anchor-grid sizes (vector and scalar paths, more row groups than rows, single cells) for the four
    directions: forward results vs the numpy oracle, backward results vs the adjoint identity <A x, g> = <x, A^T g>."""
    from mmidet_b200 import ops
    rng = np.random.default_rng(11)
    for it in range(40):
        hs, ws = int(rng.integers(1, 13)), int(rng.integers(1, 13))
        H = int(rng.integers(hs, 70))
        W = int(rng.integers(ws, 70)) if it % 3 else 4 * int(rng.integers((ws + 3) // 4, 18))
        BC = (int(rng.integers(1, 4)), int(rng.integers(1, 6)))
        x = rng.standard_normal(BC + (H, W)).astype(np.float32)
        s = rng.standard_normal(BC + (hs, ws)).astype(np.float32)
        xt, st = _t(x).requires_grad_(True), _t(s).requires_grad_(True)
        pool, up = ops.adaptive_avg_pool(xt, (hs, ws)), ops.upsample_bilinear(st, (H, W))
        tag = f"case {it}: {BC} {H}x{W} -> {hs}x{ws}"
        assert relerr(pool.detach().cpu().numpy(), O.adaptive_avg_pool2d(x, (hs, ws))) <= 1e-5, tag
        assert relerr(up.detach().cpu().numpy(), O.upsample_bilinear(s, (H, W))) <= 1e-5, tag
        (dx,) = torch.autograd.grad(pool, xt, st.detach())   # A^T s
        (ds,) = torch.autograd.grad(up, st, xt.detach())     # U^T x
        lhs_p, rhs_p = float((pool.detach().double() * st.detach().double()).sum()), float((xt.detach().double() * dx.double()).sum())
        lhs_u, rhs_u = float((up.detach().double() * xt.detach().double()).sum()), float((st.detach().double() * ds.double()).sum())
        scale_p = float(pool.detach().abs().double().sum()) + 1.0
        scale_u = float(up.detach().abs().double().sum()) + 1.0
        assert abs(lhs_p - rhs_p) <= 1e-4 * scale_p, tag
        assert abs(lhs_u - rhs_u) <= 1e-4 * scale_u, tag


def test_new_kernels_stay_inside_their_outputs():
    """no memcheck tool on the GPU pool: every output of the resampling / token-layout / pattern kernels is carved out of a
    sentinel-filled arena and the guard bands on both sides must survive (ragged shapes: partial tiles, vector + scalar paths)."""
    from mmidet_b200 import _lib, ops
    lib = _lib.load()
    P, DT, ST = ops._ptr, ops._DT, ops._stream
    GUARD = 4096

    def arena(n, dtype):
        buf = torch.full((n + 2 * GUARD,), 7.0, device="cuda", dtype=dtype)  # sentinel 7 everywhere
        return buf, buf[GUARD:GUARD + n]

    def intact(buf, n):
        return bool((buf[:GUARD] == 7).all()) and bool((buf[GUARD + n:] == 7).all())

    for dt in (torch.float32, torch.bfloat16):
        for (B, C, H, W, hs, ws) in [(2, 3, 160, 160, 8, 8), (1, 5, 37, 53, 8, 8), (3, 100, 96, 100, 8, 8), (2, 150, 128, 136, 8, 8),
                                     (1, 7, 20, 24, 4, 6)]:
            big = torch.randn(B, C, H, W, device="cuda").to(dt)
            small = torch.randn(B, C, hs, ws, device="cuda").to(dt)
            for fn, src, n in (("mmi_avgpool_fwd", big, B * C * hs * ws), ("mmi_upsample_bilinear_bwd", big, B * C * hs * ws),
                               ("mmi_upsample_bilinear_fwd", small, B * C * H * W), ("mmi_avgpool_bwd", small, B * C * H * W)):
                buf, out = arena(n, dt)
                _lib.check(getattr(lib, fn)(P(src), P(out), B * C, H, W, hs, ws, DT[dt], ST(src)), fn)
                torch.cuda.synchronize()
                assert intact(buf, n), (fn, dt, B, C, H, W)
                assert bool(torch.isfinite(out.float()).all()) and not bool((out == 7).all())
        for (B, C, HW) in [(2, 136, 168), (1, 72, 63), (3, 256, 400), (1, 8, 16), (2, 200, 272)]:
            rgb, ir = torch.randn(B, C, HW, device="cuda").to(dt), torch.randn(B, C, HW, device="cuda").to(dt)
            n = B * 2 * HW * C
            buf, tok = arena(n, dt)
            _lib.check(lib.mmi_tokens_gather(P(rgb), P(ir), P(tok), B, C, HW, DT[dt], ST(rgb)), "mmi_tokens_gather")
            b1, o1 = arena(B * C * HW, dt)
            b2, o2 = arena(B * C * HW, dt)
            _lib.check(lib.mmi_tokens_scatter(P(tok), P(o1), P(o2), B, C, HW, DT[dt], ST(rgb)), "mmi_tokens_scatter")
            torch.cuda.synchronize()
            assert intact(buf, n) and intact(b1, B * C * HW) and intact(b2, B * C * HW), (dt, B, C, HW)
            assert torch.equal(o1.view(B, C, HW), rgb) and torch.equal(o2.view(B, C, HW), ir)
    # pattern path: tokens, rows, loss and the backward outputs
    for (B, C, h, w) in [(1, 8, 8, 8), (9, 40, 5, 7), (3, 100, 8, 8)]:
        Pn = h * w
        vis, ir = torch.randn(B, C, h, w, device="cuda"), torch.randn(B, C, h, w, device="cuda")
        w1, w2 = torch.randn(8, C, device="cuda"), torch.randn(C, 8, device="cuda")
        ws = torch.empty(lib.mmi_ffm_pattern_ws_bytes(B, C, Pn), dtype=torch.uint8, device="cuda")
        bt, tok = arena(B * 2 * Pn * C, torch.float32)
        br, rows = arena(18 * B * Pn, torch.float32)
        bl, loss = arena(1, torch.float32)
        _lib.check(lib.mmi_ffm_pattern_fwd(P(vis), P(ir), P(w1), P(w2), P(tok), P(rows), P(loss), P(ws), B, C, h, w, DT[torch.float32],
                                           ST(vis)), "mmi_ffm_pattern_fwd")
        dtok = torch.randn(B, 2 * Pn, C, device="cuda")
        bv, dvis = arena(B * C * Pn, torch.float32)
        bi, dir_ = arena(B * C * Pn, torch.float32)
        b1, dw1 = arena(8 * C, torch.float32)
        b2, dw2 = arena(8 * C, torch.float32)
        _lib.check(lib.mmi_ffm_pattern_bwd(P(vis), P(ir), P(dtok), P(rows), P(w1), P(w2), P(dvis), P(dir_), P(dw1), P(dw2), P(ws), B, C, Pn,
                                           DT[torch.float32], ST(vis)), "mmi_ffm_pattern_bwd")
        torch.cuda.synchronize()
        for buf, n in ((bt, B * 2 * Pn * C), (br, 18 * B * Pn), (bl, 1), (bv, B * C * Pn), (bi, B * C * Pn), (b1, 8 * C), (b2, 8 * C)):
            assert intact(buf, n), (B, C, h, w, n)
