"""CPU tests of the plug-in helper: mmidet_b200.mamba.install binds the B200 path into a reference-shaped module tree
by name and uninstall restores it.  A stand-in tree is used (the real reference does not travel to CI)."""
import types

import torch


def _fake_tree():
    class MambaBlock:  # same attribute names the reference class exposes
        def selective_scan(self, *a):
            return "ref_scan"

        def selective_scan_seq(self, *a):
            return "ref_seq"

    class GPT1_fourier:  # models/common.py:300
        def forward(self, x):
            return "ref_fourier"

    class GPT1:  # models/common.py:140
        def forward(self, x):
            return "ref_gpt1"

    mamba = types.SimpleNamespace(MambaBlock=MambaBlock, pscan="ref_pscan")
    common = types.SimpleNamespace(extract_frequency2="ref_ffm", Seperation_loss="ref_sep", GPT1_fourier=GPT1_fourier, GPT1=GPT1)
    yolo = types.SimpleNamespace(GPT="ref_gpt")
    return types.SimpleNamespace(mamba=mamba, common=common, yolo_test=yolo)


def test_install_and_uninstall():
    from mmidet_b200 import ffm, mamba as M, pscan
    t = _fake_tree()
    saved = M.install(ref_models=t, fusion=True)
    assert t.mamba.pscan is pscan.pscan
    assert t.common.extract_frequency2 is ffm.extract_frequency2
    assert t.common.Seperation_loss is ffm.separation_loss
    assert t.common.GPT1_fourier.forward is ffm.fourier_forward
    assert t.common.GPT1.forward is ffm.gpt1_forward
    assert t.yolo_test.GPT is M.MambaFusion
    assert t.mamba.MambaBlock.selective_scan is t.mamba.MambaBlock.selective_scan_seq
    # the patched method is the fused operator: CPU tensors must raise, never fall back
    try:
        t.mamba.MambaBlock().selective_scan(torch.randn(1, 8, 16), torch.randn(1, 8, 16), torch.randn(16, 16),
                                            torch.randn(1, 8, 16), torch.randn(1, 8, 16), torch.randn(16))
        raise AssertionError("expected RuntimeError")
    except RuntimeError:
        pass
    M.uninstall(saved)
    assert t.mamba.pscan == "ref_pscan" and t.yolo_test.GPT == "ref_gpt" and t.common.extract_frequency2 == "ref_ffm"
    assert t.mamba.MambaBlock().selective_scan() == "ref_scan"
    assert t.common.GPT1_fourier().forward(None) == "ref_fourier"
    assert t.common.GPT1().forward(None) == "ref_gpt1"


def test_state_dict_layout_matches_reference(golden):
    """parameter names and shapes of ResidualBlock == the reference's (tests/golden/mamba_block.npz carries its state_dict)."""
    from mmidet_b200.mamba import MambaConfig, ResidualBlock
    g = golden("mamba_block")
    ref = {k[3:]: v.shape for k, v in g.items() if k.startswith("sd.")}
    ours = {k: tuple(v.shape) for k, v in ResidualBlock(MambaConfig(d_model=16, n_layers=1)).state_dict().items()}
    assert ours == {k: tuple(s) for k, s in ref.items()}


def test_fusion_contract_shapes_on_meta():
    """GPT contract (models/common.py:1270-1370): ctor takes d_model (+ ignored extras), token order is VIS then IR."""
    from mmidet_b200.mamba import MambaFusion
    fus = MambaFusion(32, 8, 4, 1, 8, 8, 0.1, 0.1, 0.1)
    assert fus.n_embd == 32 and len(fus.layers) == 1
    names = {k.split(".")[3] for k in fus.state_dict() if k.startswith("layers.0.mixer.")}
    assert names >= {"A_log", "D", "in_proj", "conv1d", "x_proj", "dt_proj", "out_proj"}


def test_staged_reference_is_unmodified():
    """baseline/_ref (what travels to the GPU box) is byte-for-byte the reference: every staged file equals its source
    under /root/reference and the manifest matches.  Skipped where either side is absent."""
    import hashlib
    import os
    import pytest
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dst, src = os.path.join(root, "baseline", "_ref"), os.environ.get("MMIDET_REF", "/root/reference")
    if not os.path.isdir(dst):
        pytest.skip("no staged reference")
    lines = [l.split("  ", 1) for l in open(os.path.join(dst, "MANIFEST.sha256")).read().splitlines()]
    assert len(lines) > 20
    for digest, rel in lines:
        data = open(os.path.join(dst, rel), "rb").read()
        assert hashlib.sha256(data).hexdigest() == digest, rel
        if os.path.isdir(src):
            assert data == open(os.path.join(src, rel), "rb").read(), rel


def test_checkpoint_loader_roundtrip(tmp_path):
    """harness.load_checkpoint == the reference's attempt_load steps (models/experimental.py:113-134) on a checkpoint in
    train.py:882-894's format: the pickled module comes back, fused, in eval mode, and computes the same outputs."""
    import pytest
    import torch
    from mmidet_b200 import harness as H
    try:
        ref = H.import_reference()
    except RuntimeError:
        pytest.skip("no reference checkout")
    model = H.build_detector("s", "pytorch", device="cpu").eval()
    x = torch.rand(1, 3, 64, 64)
    with torch.no_grad():
        want = model(x, x)[0][0]
    path = str(tmp_path / "last.pt")
    torch.save({"epoch": 3, "model": model, "ema": None, "optimizer": None}, path)
    got_model = H.load_checkpoint(path, install_path=False)
    assert not got_model.training
    assert not any(hasattr(m, "bn") for m in got_model.modules() if type(m) is ref.common.Conv)  # Conv + BN folded
    with torch.no_grad():
        got = got_model(x, x)[0][0]
    assert torch.allclose(got, want, atol=1e-4, rtol=1e-4)


def test_harness_optimizer_groups_follow_the_reference_quirk():
    """train.py:572-579 groups parameters by MODULE attribute: bare nn.Parameters (MambaBlock.A_log / D) land in no group
    and are never updated -- the harness keeps that, which is why A stays on the geometric fast path."""
    import pytest
    import torch
    from mmidet_b200 import harness as H
    try:
        H.import_reference()
    except RuntimeError:
        pytest.skip("no reference checkout")
    model = H.build_detector("s", "pytorch", device="cpu")
    hyp = H.scale_hyp(model, 6, 640)
    opt = H.make_optimizer(model, hyp, 16)
    in_groups = {id(p) for g in opt.param_groups for p in g["params"]}
    names = dict(model.named_parameters())
    bare = [n for n in names if n.endswith("A_log") or n.endswith(".D")]
    assert len(bare) == 8  # four fusion sites x (A_log, D)
    assert all(id(names[n]) not in in_groups for n in bare)
    assert all(id(p) in in_groups for n, p in names.items() if n.endswith("in_proj.weight") or n.endswith("conv1d.bias"))
    assert len(opt.param_groups) == 3 and opt.param_groups[1]["weight_decay"] > 0 and opt.param_groups[0]["nesterov"]
    imgs, targets = H.synthetic_batch(3, 64, device="cpu", seed=1)
    assert imgs.dtype == torch.uint8 and tuple(imgs.shape) == (3, 6, 64, 64)
    assert targets.shape[1] == 6 and float(targets[:, 0].max()) == 2.0 and float(targets[:, 2:].max()) <= 1.0
