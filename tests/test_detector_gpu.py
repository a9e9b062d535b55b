"""BASELINE configs[1] / [3] on the GPU: the UNMODIFIED reference detector (staged checkout, scripts/stage_reference.py)
with the CUDA fusion path plugged in by name (mamba.install-style binding of the YAML name GPT) against the same detector on
the reference's own pure-PyTorch MambaBlock / pscan -- both arms on the GPU, same weights."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H():
    from mmidet_b200 import harness
    try:
        harness.import_reference()
    except RuntimeError:
        pytest.skip("reference checkout not staged (run scripts/stage_reference.py where /root/reference exists)")
    return harness


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def test_detector_logits_match_reference_forward(H):
    """configs[1]: two-stream YOLOv5s, 640x640 synthetic RGB+IR pair, batch 1: Detect output (1, 25200, 11) and the three
    raw maps of the CUDA path vs the reference forward (fp32, TF32 off so that both arms run the same conv arithmetic)."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref_model = H.build_detector("s", "pytorch", seed=0).eval()
    our_model = H.build_detector("s", "ours", seed=0, state_dict=ref_model.state_dict()).eval()
    g = torch.Generator().manual_seed(2)
    rgb, ir = torch.rand(1, 3, 640, 640, generator=g).cuda(), torch.rand(1, 3, 640, 640, generator=g).cuda()
    with torch.no_grad():
        (zr, xr), _ = ref_model(rgb, ir)
        (zo, xo), _ = our_model(rgb, ir)
    assert tuple(zo.shape) == (1, 25200, 11)
    assert _rel(zo, zr) <= 1e-4
    for a, b in zip(xo, xr):
        assert _rel(a, b) <= 1e-4


def test_install_binds_into_the_real_reference_modules(H):
    """mamba.install() on the real models.mamba / models.common / models.yolo_test: the reference's own
    MambaBlock then runs the fused kernel and matches its pure-PyTorch self; YAML rows naming GPT build MambaFusion."""
    from mmidet_b200 import mamba
    ref = H.import_reference()
    torch.manual_seed(0)
    blk = ref.mamba.ResidualBlock(ref.mamba.MambaConfig(d_model=32, n_layers=1)).cuda()
    x = torch.randn(2, 300, 32, device="cuda", requires_grad=True)
    y0 = blk(x)
    (g0,) = torch.autograd.grad(y0.sum(), x)
    saved = mamba.install(fusion=True)
    try:
        assert ref.yolo_test.GPT is mamba.MambaFusion
        y1 = blk(x)
        (g1,) = torch.autograd.grad(y1.sum(), x)
        A = torch.rand(2, 100, 8, 16, device="cuda") * 0.5 + 0.5
        X = torch.randn(2, 100, 8, 16, device="cuda")
        h1 = ref.mamba.pscan(A, X)
    finally:
        mamba.uninstall(saved)
    h0 = ref.mamba.pscan(A, X)
    assert _rel(y1, y0) <= 1e-4 and _rel(g1, g0) <= 1e-4 and _rel(h1, h0) <= 1e-5


def test_training_step_both_arms(H):
    """configs[3] in miniature (YOLOv5s, 320 px, 2 pairs, fp32): one training step of the unmodified loop pieces (ComputeLoss
    unchanged) on both arms from the same weights: same loss, finite gradients, the optimizer moves the weights."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = H.import_reference()
    losses = {}
    sd = None
    for arm in ("pytorch", "ours"):
        model = H.build_detector("s", arm, seed=0, state_dict=sd).train()
        sd = sd or {k: v.clone() for k, v in model.state_dict().items()}
        hyp = H.scale_hyp(model, 6, 320)
        cl = ref.loss.ComputeLoss(model)
        opt = H.make_optimizer(model, hyp, 2)
        imgs, targets = H.synthetic_batch(2, 320, seed=5)
        w0 = model.model[0].conv.conv.weight.detach().clone() if hasattr(model.model[0], "conv") else None
        losses[arm] = float(H.train_step(model, cl, opt, imgs, targets, autocast_dtype=None, fused_prep=arm == "ours"))
        assert torch.isfinite(torch.tensor(losses[arm]))
        if w0 is not None:
            assert not torch.equal(w0, model.model[0].conv.conv.weight.detach())
    assert abs(losses["ours"] - losses["pytorch"]) <= 1e-3 * abs(losses["pytorch"])


def test_training_step_fp16_with_grad_scaler(H):
    """the reference's own precision recipe (train.py:706, :784-801): fp16 autocast + GradScaler.  Three steps on the CUDA
    fusion path: finite loss every step, the scaler stays positive and the weights move once a step is not skipped."""
    ref = H.import_reference()
    model = H.build_detector("s", "ours", seed=0, channels_last=True).train()
    hyp = H.scale_hyp(model, 6, 320)
    cl = ref.loss.ComputeLoss(model)
    opt = H.make_optimizer(model, hyp, 2)
    scaler = H.make_scaler(torch.float16)
    assert scaler is not None and H.make_scaler(torch.bfloat16) is None and H.make_scaler(None) is None
    imgs, targets = H.synthetic_batch(2, 320, seed=6)
    w0 = [p.detach().clone() for p in model.parameters()][:8]
    for _ in range(3):
        loss = H.train_step(model, cl, opt, imgs, targets, autocast_dtype=torch.float16, scaler=scaler)
        assert bool(torch.isfinite(loss))
    assert scaler.get_scale() > 0
    assert any(not torch.equal(a, b.detach()) for a, b in zip(w0, model.parameters()))


def test_quiet_removes_the_per_step_prints(H, capsys):
    """the reference prints CUDA tensors (a device sync each) from forward_once and ComputeLoss; quiet()
    shadows `print` in those modules without touching their source, and un-quieting restores it."""
    ref = H.import_reference(quiet=True)
    model = H.build_detector("s", "ours", seed=0).eval()
    x = torch.rand(1, 3, 64, 64, device="cuda")
    capsys.readouterr()
    with torch.no_grad():
        model(x, x)
    assert capsys.readouterr().out == ""
    H.quiet(ref, on=False)
    try:
        with torch.no_grad():
            model(x, x)
        assert "Combine_loss" in capsys.readouterr().out
    finally:
        H.quiet(ref, on=True)
