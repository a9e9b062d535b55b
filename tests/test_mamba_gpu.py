"""GPU parity of the host-side module layer (mmidet_b200.mamba) against outputs of the UNMODIFIED reference modules
(tests/golden/make_golden.py: gen_block, gen_fusion, gen_detector).  fp32 tolerance 1e-4 (north_star), measured as
max|a-b| / max|b|."""
import numpy as np
import pytest
import torch

from tests.util import relerr

pytestmark = pytest.mark.gpu
TOL32 = 1e-4


def _load_sd(module, g, prefix):
    sd = {k[len(prefix):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(prefix)}
    missing, unexpected = module.load_state_dict(sd, strict=True)
    assert not missing and not unexpected


def test_residual_block_matches_reference(golden):
    """reference ResidualBlock state_dict loads unchanged; forward, input gradient and every parameter gradient match."""
    from mmidet_b200.mamba import MambaConfig, ResidualBlock
    g = golden("mamba_block")
    blk = ResidualBlock(MambaConfig(d_model=16, n_layers=1))
    _load_sd(blk, g, "sd.")
    blk = blk.cuda()
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    y_mixer = blk.mixer(x)
    y = blk(x)
    assert relerr(y_mixer.detach().cpu().numpy(), g["y_mixer"]) <= TOL32
    assert relerr(y.detach().cpu().numpy(), g["y_res"]) <= TOL32
    grads = torch.autograd.grad(y, [x] + list(blk.parameters()), torch.from_numpy(g["g"]).cuda())
    assert relerr(grads[0].cpu().numpy(), g["gx"]) <= TOL32
    for (name, _), gr in zip(blk.named_parameters(), grads[1:]):
        assert relerr(gr.cpu().numpy(), g["pg." + name]) <= 2e-4, name


def test_fusion_block_matches_reference(golden):
    """MambaFusion on our kernels == the same wrapper on the reference's pure-PyTorch blocks (outputs and input grads)."""
    from mmidet_b200.mamba import MambaFusion
    g = golden("fusion_block")
    fus = MambaFusion(16, n_layer=2)
    _load_sd(fus, g, "sd.")
    fus = fus.cuda()
    rgb = torch.from_numpy(g["rgb"]).cuda().requires_grad_(True)
    ir = torch.from_numpy(g["ir"]).cuda().requires_grad_(True)
    o_rgb, o_ir = fus([rgb, ir])
    assert o_rgb.shape == rgb.shape and o_ir.shape == ir.shape
    assert relerr(o_rgb.detach().cpu().numpy(), g["o_rgb"]) <= TOL32
    assert relerr(o_ir.detach().cpu().numpy(), g["o_ir"]) <= TOL32
    d_rgb, d_ir = torch.autograd.grad([o_rgb, o_ir], [rgb, ir], [torch.from_numpy(g["g_rgb"]).cuda(), torch.from_numpy(g["g_ir"]).cuda()])
    assert relerr(d_rgb.cpu().numpy(), g["d_rgb"]) <= TOL32
    assert relerr(d_ir.cpu().numpy(), g["d_ir"]) <= TOL32


def test_detector_fusion_sites_match_reference(golden):
    """BASELINE configs[1] at the fusion boundary: the feature maps entering and leaving the four fusion call sites of the
    UNMODIFIED two-stream YOLOv5s (models/yolo_test.py Model + YAML, GPT bound to MambaFusion on reference blocks).  The
    block weights are rebuilt from the per-site seed; the checksums in the fixture prove they are the reference's."""
    from mmidet_b200.mamba import MambaFusion
    g = golden("detector_fusion")
    for i in range(int(g["n_sites"])):
        d_model = int(g[f"f{i}.d_model"])
        torch.manual_seed(1000 + i)
        fus = MambaFusion(d_model, n_layer=1)
        wsum = np.array([[float(v.double().sum()), float(v.double().abs().sum())] for _, v in sorted(fus.state_dict().items())])
        if not np.allclose(wsum, g[f"f{i}.wsum"], rtol=1e-9, atol=1e-9):
            pytest.fail("seeded init does not reproduce the fixture's weights (torch RNG drift): regenerate "
                        "tests/golden/detector_fusion.npz with make_golden.py detector")
        fus = fus.cuda().eval()
        with torch.no_grad():
            o_rgb, o_ir = fus([torch.from_numpy(g[f"f{i}.rgb"]).cuda(), torch.from_numpy(g[f"f{i}.ir"]).cuda()])
        assert relerr(o_rgb.cpu().numpy(), g[f"f{i}.o_rgb"]) <= TOL32, (i, d_model)
        assert relerr(o_ir.cpu().numpy(), g[f"f{i}.o_ir"]) <= TOL32, (i, d_model)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_block_is_half_clean(dtype):
    """the reference block breaks under .half()/.bfloat16(); ours must run and stay within the bf16 budget
    of the fp32 block on the same (rounded) weights."""
    from mmidet_b200.mamba import MambaBlock, MambaConfig
    torch.manual_seed(3)
    blk = MambaBlock(MambaConfig(d_model=32, n_layers=1)).cuda()
    x = torch.randn(2, 100, 32, device="cuda")
    ref = blk(x).detach()
    low = blk.to(dtype)
    assert low.A_log.dtype == torch.float32 and low.D.dtype == torch.float32  # fp32 masters survive .half() / .bfloat16()
    assert low.in_proj.weight.dtype == dtype
    out = low(x.to(dtype))
    assert out.dtype == dtype
    assert relerr(out.detach().float().cpu().numpy(), ref.cpu().numpy()) <= 3e-2


def test_module_deepcopy_and_pickle():
    """ModelEMA deep-copies the model (utils/torch_utils.py:281) and checkpoints pickle it (train.py:885)."""
    import copy
    import pickle
    from mmidet_b200.mamba import MambaFusion
    fus = MambaFusion(16).cuda()
    x = [torch.randn(1, 16, 4, 4, device="cuda"), torch.randn(1, 16, 4, 4, device="cuda")]
    a = fus(x)
    b = copy.deepcopy(fus)(x)
    c = pickle.loads(pickle.dumps(fus))(x)
    for u, v, w in zip(a, b, c):
        assert torch.equal(u, v) and torch.equal(u, w)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("shape,K", [((2, 70, 40), 4), ((1, 3, 8), 4), ((3, 129, 136), 3), ((2, 257, 512), 4), ((1, 33, 16), 2)])
def test_causal_conv1d_silu_vs_oracle(shape, K, dtype, tol):
    """channels-last depthwise causal conv + SiLU, forward and all three gradients, incl. a strided (chunk view) input,
    segment-boundary lengths (33, 129, 257 = kConvSeg multiples + 1) and ED not a multiple of the 128-channel block."""
    from mmidet_b200 import ops
    from oracle import oracle as O
    B, L, ED = shape
    rng = np.random.default_rng(L + ED)
    xz = torch.from_numpy(rng.standard_normal((B, L, 2 * ED)).astype(np.float32)).cuda().to(dtype).requires_grad_(True)
    w = torch.from_numpy((rng.standard_normal((ED, 1, K)) * 0.5).astype(np.float32)).cuda().requires_grad_(True)
    b = torch.from_numpy(rng.standard_normal(ED).astype(np.float32)).cuda().requires_grad_(True)
    g = torch.from_numpy(rng.standard_normal((B, L, ED)).astype(np.float32)).cuda().to(dtype)
    x = xz.chunk(2, dim=-1)[0]  # row pitch 2*ED, as in MambaBlock.forward
    y = ops.causal_conv1d_silu(x, w, b)
    gxz, gw, gb = torch.autograd.grad(y, [xz, w, b], g)
    xn = x.detach().float().cpu().numpy()
    yr = O.causal_conv1d_silu(xn, w.detach().cpu().numpy()[:, 0, :], b.detach().cpu().numpy())
    dx, dw, db = O.causal_conv1d_silu_bwd(xn, w.detach().cpu().numpy()[:, 0, :], b.detach().cpu().numpy(), g.float().cpu().numpy())
    assert relerr(y.detach().float().cpu().numpy(), yr) <= tol
    assert relerr(gxz[..., :ED].float().cpu().numpy(), dx) <= tol
    assert float(gxz[..., ED:].abs().max()) == 0.0
    assert relerr(gw.cpu().numpy()[:, 0, :], dw) <= max(tol, 2e-5)
    assert relerr(gb.cpu().numpy(), db) <= max(tol, 2e-5)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("rows,C", [(7, 8), (130, 256), (33, 1000), (5, 1024), (19, 1280), (9, 2560), (3, 4096), (11, 1032)])
def test_rmsnorm_vs_torch_fp64(rows, C, dtype, tol):
    """fused RMSNorm (models/mamba.py:356-366) forward, dx and dw vs the reference formula evaluated in fp64."""
    from mmidet_b200 import ops
    torch.manual_seed(rows + C)
    x = torch.randn(3, rows, C, device="cuda").to(dtype).requires_grad_(True)
    w = (torch.rand(C, device="cuda") + 0.5).requires_grad_(True)
    g = torch.randn(3, rows, C, device="cuda").to(dtype)
    y = ops.rmsnorm(x, w, 1e-5)
    gx, gw = torch.autograd.grad(y, [x, w], g)
    xr = x.detach().double().requires_grad_(True)
    wr = w.detach().double().requires_grad_(True)
    yr = xr * torch.rsqrt(xr.pow(2).mean(-1, keepdim=True) + 1e-5) * wr
    gxr, gwr = torch.autograd.grad(yr, [xr, wr], g.double())
    assert relerr(y.detach().float().cpu().numpy(), yr.detach().cpu().numpy()) <= tol
    assert relerr(gx.float().cpu().numpy(), gxr.cpu().numpy()) <= tol
    assert relerr(gw.float().cpu().numpy(), gwr.cpu().numpy()) <= max(tol, 5e-5)


@pytest.mark.parametrize("ac", [torch.bfloat16, torch.float16])
def test_rmsnorm_under_autocast_emits_gemm_dtype(ac):
    """under autocast an fp32 input gives the 16-bit tensor the following GEMM would cast to -- bit-identical to the fp32
    result followed by .to(ac) -- and the backward takes the 16-bit gradient directly (dx, dw in fp32)."""
    from mmidet_b200 import ops
    torch.manual_seed(9)
    rows, C = 300, 256
    x = torch.randn(2, rows, C, device="cuda", requires_grad=True)
    w = (torch.rand(C, device="cuda") + 0.5).requires_grad_(True)
    g = torch.randn(2, rows, C, device="cuda").to(ac)
    y32 = ops.rmsnorm(x, w, 1e-5)
    with torch.autocast("cuda", dtype=ac):
        y = ops.rmsnorm(x, w, 1e-5)
    assert y.dtype == ac and y32.dtype == torch.float32
    assert torch.equal(y, y32.detach().to(ac))
    gx, gw = torch.autograd.grad(y, [x, w], g)
    gx32, gw32 = torch.autograd.grad(y32, [x, w], g.float())
    assert gx.dtype == torch.float32 and gw.dtype == torch.float32
    assert relerr(gx.cpu().numpy(), gx32.cpu().numpy()) <= 1e-5
    assert relerr(gw.cpu().numpy(), gw32.cpu().numpy()) <= 5e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 16, 6, 5), (1, 72, 7, 9), (3, 256, 20, 20), (2, 136, 12, 14), (1, 8, 4, 4),
                                   (2, 64, 80, 80), (1, 200, 16, 17)])  # 16-byte-vector path with partial tiles, and scalar path
def test_tokens_gather_scatter(shape, dtype):
    """token layout kernels == flatten / cat / transpose of models/common.py:1338-1343 (bit-exact: pure data movement), the
    scatter inverts the gather, and each is the other's adjoint under autograd."""
    from mmidet_b200 import ops
    B, C, H, W = shape
    torch.manual_seed(C)
    rgb = torch.randn(shape, device="cuda").to(dtype).requires_grad_(True)
    ir = torch.randn(shape, device="cuda").to(dtype).requires_grad_(True)
    tok = ops.tokens_gather(rgb, ir)
    ref = torch.cat([rgb.flatten(2), ir.flatten(2)], dim=2).transpose(1, 2)
    assert torch.equal(tok, ref)
    r2, i2 = ops.tokens_scatter(tok, shape)
    assert torch.equal(r2, rgb) and torch.equal(i2, ir)
    g = torch.randn_like(tok)
    gr, gi = torch.autograd.grad(tok, [rgb, ir], g)
    gr_ref, gi_ref = torch.autograd.grad(ref, [rgb, ir], g)
    assert torch.equal(gr, gr_ref) and torch.equal(gi, gi_ref)


def test_graphed_fusion_matches_eager():
    """CUDA-graph replay of a fusion block (inference): same outputs as the eager call, for two different inputs of the same
    shape (replay reads the new data) and for a second shape (second graph)."""
    from mmidet_b200.graphs import Graphed
    from mmidet_b200.mamba import MambaFusion
    torch.manual_seed(0)
    fus = MambaFusion(64).cuda().eval()
    fast = Graphed(fus)
    for shape in ((1, 64, 12, 12), (1, 64, 12, 12), (2, 64, 7, 9)):
        x = [torch.randn(shape, device="cuda"), torch.randn(shape, device="cuda")]
        with torch.no_grad():
            a, b = fus(x)
        c, d = fast(x)
        assert torch.allclose(a, c, atol=1e-6, rtol=1e-5) and torch.allclose(b, d, atol=1e-6, rtol=1e-5)
    assert len(fast._graphs) == 2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conv_and_rmsnorm_stay_inside_their_outputs(dtype):
    """outputs of the conv prologue and RMSNorm kernels carved out of sentinel arenas (ragged L, ED / C that do not fill a
    128-channel group, rows that do not fill a warp strip): guard bands must survive -- stands in for a memcheck tool."""
    from mmidet_b200 import _lib, ops
    lib = _lib.load()
    P, DT, ST = ops._ptr, ops._DT, ops._stream
    GUARD = 4096
    arenas = []

    def carve(n, dt):
        buf = torch.full((n + 2 * GUARD,), 7.0, device="cuda", dtype=dt)
        arenas.append((buf, n))
        return buf[GUARD:GUARD + n]

    for (B, L, ED) in [(2, 77, 72), (1, 130, 8), (3, 33, 136), (1, 257, 512)]:
        x, gy = torch.randn(B, L, ED, device="cuda").to(dtype), torch.randn(B, L, ED, device="cuda").to(dtype)
        w, b = torch.randn(ED, 4, device="cuda"), torch.randn(ED, device="cuda")
        y, dx = carve(B * L * ED, dtype), carve(B * L * ED, dtype)
        dw, db = carve(ED * 4, torch.float32), carve(ED, torch.float32)
        _lib.check(lib.mmi_causal_conv1d_fwd(P(x), P(w), P(b), P(y), B, L, ED, 4, ED, ED, DT[dtype], 1, ST(x)), "conv fwd")
        _lib.check(lib.mmi_causal_conv1d_bwd(P(x), P(w), P(b), P(gy), P(dx), P(dw), P(db), B, L, ED, 4, ED, ED, ED, DT[dtype], 1, ST(x)),
                   "conv bwd")
    for (rows, C) in [(77, 72), (1, 8), (4097, 256), (130, 1024), (21, 1280), (9, 2568)]:
        x, g = torch.randn(rows, C, device="cuda").to(dtype), torch.randn(rows, C, device="cuda").to(dtype)
        w = torch.rand(C, device="cuda") + 0.5
        y, dx, dw = carve(rows * C, dtype), carve(rows * C, dtype), carve(C, torch.float32)
        _lib.check(lib.mmi_rmsnorm_fwd(P(x), P(w), P(y), rows, C, C, C, 1e-5, DT[dtype], -1, ST(x)), "rmsnorm fwd")
        _lib.check(lib.mmi_rmsnorm_bwd(P(x), P(w), P(g), P(dx), P(dw), rows, C, C, C, C, 1e-5, DT[dtype], -1, ST(x)), "rmsnorm bwd")
    torch.cuda.synchronize()
    for i, (buf, n) in enumerate(arenas):
        assert bool((buf[:GUARD] == 7).all()) and bool((buf[GUARD + n:] == 7).all()), f"guard band of arena {i} overwritten"
        assert bool(torch.isfinite(buf[GUARD:GUARD + n].float()).all())


def test_step_matches_reference_and_forward(golden):
    """MambaBlock.step / ssm_step (models/mamba.py:289-353) through ResidualBlock.step over 8 tokens from the empty cache
    (h = None, zero conv window): outputs, the state after every token and the final conv window against the unmodified
    reference, and against our own forward() on the same prefix (recurrent form == parallel form).  General (trained) A."""
    from mmidet_b200.mamba import MambaConfig, ResidualBlock
    g = golden("mamba_step")
    cfg = MambaConfig(d_model=16, n_layers=1)
    blk = ResidualBlock(cfg)
    _load_sd(blk, g, "sd.")
    blk = blk.cuda().eval()
    x = torch.from_numpy(g["x"]).cuda()
    B, T, _ = x.shape
    cache = (None, torch.zeros(B, cfg.d_inner, cfg.d_conv - 1, device="cuda"))
    ys, hs = [], []
    with torch.no_grad():
        for t in range(T):
            y, cache = blk.step(x[:, t], cache)
            ys.append(y)
            hs.append(cache[0])
        y_fwd = blk(x)
    y_step = torch.stack(ys, 1).cpu().numpy()
    assert relerr(y_step, g["y_step"]) <= TOL32
    assert relerr(torch.stack(hs, 1).cpu().numpy(), g["h_step"]) <= TOL32
    assert relerr(cache[1].cpu().numpy(), g["inputs_last"]) <= 1e-6
    assert relerr(y_fwd.cpu().numpy(), g["y_fwd"]) <= TOL32
    assert relerr(y_step, y_fwd.cpu().numpy()) <= TOL32


def test_no_eager_fallbacks():
    """shapes outside the kernels raise instead of running stock torch: d_conv > 4, an RMSNorm width that is not a multiple
    of 8, CPU tensors through the fusion block."""
    from mmidet_b200.mamba import MambaBlock, MambaConfig, MambaFusion, RMSNorm
    blk = MambaBlock(MambaConfig(d_model=16, n_layers=1, d_conv=5)).cuda()
    with pytest.raises(RuntimeError):
        blk(torch.randn(1, 8, 16, device="cuda"))
    with pytest.raises(RuntimeError):
        RMSNorm(12).cuda()(torch.randn(2, 3, 12, device="cuda"))
    with pytest.raises(RuntimeError):
        MambaFusion(16)([torch.randn(1, 16, 4, 4), torch.randn(1, 16, 4, 4)])


def test_inference_skips_checkpoints(monkeypatch):
    """under torch.no_grad() (eval / Graphed inference) the forward neither writes checkpoints nor saves tensors, even though
    A_log / D are Parameters with requires_grad=True; with grad enabled it does both."""
    from mmidet_b200 import ops
    B, L, ED, N = 2, 320, 64, 16
    x, delta = torch.randn(B, L, ED, device="cuda"), torch.rand(B, L, ED, device="cuda") * 0.1
    A = torch.nn.Parameter(-torch.arange(1, N + 1, device="cuda", dtype=torch.float32).repeat(ED, 1))
    D = torch.nn.Parameter(torch.ones(ED, device="cuda"))
    Bm, Cm = torch.randn(B, L, N, device="cuda"), torch.randn(B, L, N, device="cuda")
    seen = []
    real = ops.selscan_fwd_raw

    def spy(*a, **k):
        seen.append(k.get("want_chk"))
        return real(*a, **k)

    monkeypatch.setattr(ops, "selscan_fwd_raw", spy)
    with torch.no_grad():
        y = ops.selective_scan(x, delta, A, Bm, Cm, D)
    assert y.grad_fn is None and seen == [False]
    y2 = ops.selective_scan(x, delta, A, Bm, Cm, D)
    assert y2.grad_fn is not None and seen == [False, True]
    assert torch.equal(y, y2.detach())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fusion_accepts_channels_last_maps(dtype):
    """maps from a channels_last backbone take the no-transpose route (concatenate token rows, return views): same values
    and gradients as the NCHW route."""
    from mmidet_b200.mamba import MambaFusion
    torch.manual_seed(0)
    fus = MambaFusion(32).cuda().to(dtype)
    a, b = torch.randn(2, 32, 9, 7, device="cuda", dtype=dtype), torch.randn(2, 32, 9, 7, device="cuda", dtype=dtype)
    x1 = [a.clone().requires_grad_(True), b.clone().requires_grad_(True)]
    x2 = [a.clone().contiguous(memory_format=torch.channels_last).requires_grad_(True),
          b.clone().contiguous(memory_format=torch.channels_last).requires_grad_(True)]
    o1, o2 = fus(x1), fus(x2)
    assert o2[0].stride(1) == 1 and o2[0].shape == a.shape  # channel-innermost views of the token tensor (no copy)
    g = torch.randn_like(a)
    g1 = torch.autograd.grad([o1[0], o1[1]], x1, [g, g])
    g2 = torch.autograd.grad([o2[0], o2[1]], x2, [g, g])
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    for u, v in zip(o1 + g1, o2 + g2):
        assert relerr(v.detach().float().cpu().numpy(), u.detach().float().cpu().numpy()) <= tol
