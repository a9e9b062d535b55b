"""Detector harness: run the UNMODIFIED reference two-stream YOLOv5 (models/yolo_test.py Model / parse_model / YAML,
utils/loss.py ComputeLoss, utils/general.py non_max_suppression) with the B200 fusion path plugged in through
`mamba.install` -- or, for the comparison arm, with the same fusion module built on the reference's own pure-PyTorch
`models.mamba.ResidualBlock` / `models.pscan` ("the unmodified PyTorch GPU path" of the north_star).

Nothing of the reference is copied or edited: it is imported from an external checkout located by `locate_reference()`
(env MMIDET_REF, /root/reference in the build container, or the byte-for-byte staged copy baseline/_ref that
scripts/stage_reference.py makes for the GPU box).  What this module adds around it is what
the reference's own drivers do inline:

    synthetic_batch / prep_inputs     train.py:741-745 (uint8 batch -> float / 255 -> RGB | IR split), SURVEY 8d config 4
    make_optimizer / scale_hyp        train.py:567-587, :688-696 (SGD groups by module attribute; bare Parameters such as
                                      A_log / D land in no group, SURVEY App. B -- kept, it is the reference's behaviour)
    training_backend_flags            train.py:66 (cudnn.benchmark as the reference's init_seeds sets it)
    train_step / make_scaler          train.py:783-804, :706 (fp16 autocast forward, ComputeLoss, scaled backward, optimizer step)
    infer                             detect_twostream.py:88-94 (model forward + non_max_suppression timing window)
    load_checkpoint                   models/experimental.py:113-134 (attempt_load) for the checkpoints train.py:882-894 writes
    quiet()                           the per-step print()/sync points of utils/loss.py:162-182 and
                                      models/yolo_test.py:253,269: formatting a CUDA tensor for print is a
                                      device synchronisation; the module-global name `print` is shadowed by a no-op in those
                                      modules' namespaces (no source edit)."""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
import types
import unittest.mock as mock

import torch
import torch.nn as nn

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIZES = {"s": (0.33, 0.50), "m": (0.67, 0.75), "l": (1.0, 1.0), "x": (1.33, 1.25)}  # depth_multiple, width_multiple
FUSION_YAML = "models/transformer/yolov5l_fusion_transformer_M3FD.yaml"


def locate_reference() -> str:
    """The reference checkout: $MMIDET_REF, /root/reference, or the staged copy baseline/_ref.  Raises if none exists."""
    for p in (os.environ.get("MMIDET_REF"), "/root/reference", os.path.join(_ROOT, "baseline", "_ref")):
        if p and os.path.isfile(os.path.join(p, "models", "yolo_test.py")):
            return p
    raise RuntimeError("reference checkout not found: set MMIDET_REF, or run scripts/stage_reference.py where /root/reference "
                       "exists so that baseline/_ref travels with the repo")


_REF = None


def import_reference(quiet: bool = True) -> types.SimpleNamespace:
    """Import the reference's modules as black boxes: packages absent from this image that the reference
    imports at module level but never needs on this path are stubbed.  Returns a namespace of the imported modules."""
    global _REF
    if _REF is None:
        ref = locate_reference()
        for n in ("matplotlib", "matplotlib.pyplot", "seaborn", "thop", "torchsummary"):
            try:
                importlib.import_module(n)
            except Exception:
                sys.modules[n] = mock.MagicMock()
        if ref not in sys.path:
            sys.path.insert(0, ref)
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            import models.common as C
            import models.mamba as M
            import models.pscan as P
            import models.yolo_test as Y
            import utils.general as G
            import utils.loss as LS
        _REF = types.SimpleNamespace(root=ref, yolo_test=Y, common=C, mamba=M, pscan=P, loss=LS, general=G)
    if quiet:
        globals()["quiet"](_REF)
    return _REF


def quiet(ref=None, on: bool = True):
    """Shadow `print` in the reference modules that print CUDA tensors every step (each one a device sync)."""
    ref = ref or import_reference(quiet=False)
    for m in (ref.yolo_test, ref.loss, ref.common):
        if on:
            m.print = lambda *a, **k: None
        elif "print" in vars(m):
            del m.print


def detector_cfg(size: str = "l", nc: int = 6, yaml_rel: str = FUSION_YAML) -> dict:
    """The reference's fusion YAML with the depth / width multiples of YOLOv5 s / m / l / x (GPT scales with
    ch[f]; dict cfg accepted by Model, models/yolo_test.py:82-83)."""
    import yaml
    ref = import_reference()
    with open(os.path.join(ref.root, yaml_rel)) as f:
        cfg = yaml.safe_load(f)
    cfg["depth_multiple"], cfg["width_multiple"] = SIZES[size]
    cfg["nc"] = nc
    return cfg


def build_detector(size: str = "l", arm: str = "ours", nc: int = 6, seed: int = 0, n_layer: int = 1, device="cuda",
                   state_dict=None, channels_last: bool = False):
    """Model(cfg) of the unmodified reference with the YAML name `GPT` bound to
        arm="ours"     mmidet_b200.mamba.MambaFusion (fused sm_100a kernels)
        arm="pytorch"  the same MambaFusion wrapper on the reference's models.mamba.ResidualBlock (pure PyTorch pscan)
    Same seed -> same weights in both arms (or pass `state_dict` of the other arm)."""
    from . import mamba as ours
    ref = import_reference()
    Y = ref.yolo_test
    if arm == "ours":
        fusion = lambda d_model, *a, **k: ours.MambaFusion(d_model, n_layer=n_layer)  # noqa: E731
    elif arm == "pytorch":
        fusion = lambda d_model, *a, **k: ours.MambaFusion(d_model, n_layer=n_layer, block_cls=ref.mamba.ResidualBlock,  # noqa: E731
                                                           config_cls=ref.mamba.MambaConfig)
    else:
        raise ValueError(arm)
    old = Y.GPT
    Y.GPT = fusion  # parse_model eval()s the YAML name and tests `m is GPT` against this same global (yolo_test.py:560,600)
    try:
        torch.manual_seed(seed)
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            model = Y.Model(detector_cfg(size, nc), ch=3, nc=nc)
    finally:
        Y.GPT = old
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    model = model.to(device)
    if channels_last:
        # stock PyTorch switch (no reference edit): cuDNN convolutions and batch norms run NHWC -- their native layout on
        # tensor cores -- and the feature maps reach the fusion blocks as ready-made token rows
        model = model.to(memory_format=torch.channels_last)
    return model


def scale_hyp(model, nc: int, imgsz: int):
    """train.py:688-696: hyper-parameters attached to the model, scaled to layers / classes / image size."""
    import yaml
    ref = import_reference()
    with open(os.path.join(ref.root, "data", "hyp.scratch.yaml")) as f:
        hyp = yaml.safe_load(f)
    nl = model.model[-1].nl
    hyp["box"] *= 3.0 / nl
    hyp["cls"] *= nc / 80.0 * 3.0 / nl
    hyp["obj"] *= (imgsz / 640) ** 2 * 3.0 / nl
    hyp["label_smoothing"] = 0.0
    model.nc, model.hyp, model.gr = nc, hyp, 1.0
    return hyp


def make_optimizer(model, hyp, total_batch: int):
    """train.py:567-587: SGD(nesterov) with the three parameter groups picked by module attribute."""
    nbs = 64
    accumulate = max(round(nbs / total_batch), 1)
    wd = hyp["weight_decay"] * total_batch * accumulate / nbs
    pg0, pg1, pg2 = [], [], []
    for _, v in model.named_modules():
        if hasattr(v, "bias") and isinstance(v.bias, nn.Parameter):
            pg2.append(v.bias)
        if isinstance(v, nn.BatchNorm2d):
            pg0.append(v.weight)
        elif hasattr(v, "weight") and isinstance(v.weight, nn.Parameter):
            pg1.append(v.weight)
    opt = torch.optim.SGD(pg0, lr=hyp["lr0"], momentum=hyp["momentum"], nesterov=True)
    opt.add_param_group({"params": pg1, "weight_decay": wd})
    opt.add_param_group({"params": pg2})
    return opt


def synthetic_batch(B: int, imgsz: int, nc: int = 6, boxes_per_image: int = 8, seed: int = 0, device="cuda"):
    """imgs uint8 (B, 6, H, W) [RGB | IR stacked on the channel axis, utils/datasets.py], targets
    (n, 6) = (image index, class, cx, cy, w, h) normalised; fixed seed per rank."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    imgs = torch.randint(0, 256, (B, 6, imgsz, imgsz), dtype=torch.uint8, generator=g)
    n = B * boxes_per_image
    t = torch.empty(n, 6)
    t[:, 0] = torch.arange(B).repeat_interleave(boxes_per_image).float()
    t[:, 1] = torch.randint(0, nc, (n,), generator=g).float()
    t[:, 2:4] = torch.rand(n, 2, generator=g) * 0.8 + 0.1
    t[:, 4:6] = torch.rand(n, 2, generator=g) * 0.25 + 0.05
    return imgs.to(device), t.to(device)


def prepare_inference(model, dtype=torch.float16, fuse: bool = True, channels_last: bool = False):
    """What the reference's loader does before detect_twostream.py's loop: Conv + BatchNorm folding (models/experimental.py:119
    `.fuse().eval()`, the reference's own Model.fuse) and the half() of detect_twostream.py:45."""
    model = model.eval()
    if fuse:
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            model = model.fuse()  # (before any layout change: fuse_conv_and_bn views the weights as NCHW-contiguous)
    model = model.to(dtype)
    return model.to(memory_format=torch.channels_last) if channels_last else model


def load_checkpoint(path: str, map_location="cpu", fuse: bool = True, install_path: bool = True):
    """models/experimental.py:113-134 (`attempt_load`, single model) and train.py:882-894's format: a pickled dict whose
    'ema' or 'model' entry is the whole module.  The reference's own call breaks on torch >= 2.6 (`weights_only` defaults to
    True) and would try to download a missing file; this does the same steps without either: unpickle with the
    reference checkout importable, take the EMA model if present, `.float()`, the reference's `fuse()`, `.eval()`, the
    compatibility fix-ups.  install_path=True additionally binds the CUDA path onto the reference classes the checkpoint's
    modules are instances of (mamba.install: scan / pscan / FFM), so the loaded detector runs on the kernels as is."""
    ref = import_reference()
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    model = ckpt["ema" if ckpt.get("ema") else "model"].float()
    if fuse:
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            model = model.fuse()
    model = model.eval()
    for m in model.modules():
        if type(m) in (nn.Hardswish, nn.LeakyReLU, nn.ReLU, nn.ReLU6, nn.SiLU):
            m.inplace = True
        elif type(m) is ref.common.Conv:
            m._non_persistent_buffers_set = set()
    if install_path:
        from . import mamba as ours
        ours.install(scan=True, pscan=True, ffm=True, fusion=False)
    return model


def prep_inputs(imgs_u8: torch.Tensor, dtype=torch.float32):
    """train.py:743-745 / detect_twostream.py:74-85: uint8 -> float / 255, split into the RGB and IR streams -- one pass of
    the fused kernel (csrc/detect.cu) instead of .float(), / 255 and two strided slice copies."""
    from . import postprocess
    return postprocess.split_normalize(imgs_u8, dtype)


def training_backend_flags():
    """train.py:66 -> utils/general.py init_seeds(2 + rank) -> utils/torch_utils.py:42-45: with a non-zero seed the reference
    trains with `cudnn.benchmark = True, cudnn.deterministic = False` (cuDNN picks each convolution's algorithm by timing)."""
    torch.backends.cudnn.benchmark, torch.backends.cudnn.deterministic = True, False


def make_scaler(autocast_dtype=torch.float16):
    """train.py:706: `amp.GradScaler(enabled=cuda)` -- the reference trains under fp16 autocast with loss scaling.  Returns None
    for the precisions that need no scaling (bf16 autocast, fp32)."""
    return torch.amp.GradScaler("cuda") if autocast_dtype == torch.float16 else None


def train_step(model, compute_loss, optimizer, imgs_u8, targets, autocast_dtype=torch.float16, world_size: int = 1,
               fused_prep: bool = True, scaler=None):
    """train.py:783-804 for one batch: autocast forward (fp16 in the reference, :784), ComputeLoss, `scaler.scale(loss).backward()`
    (:796), `scaler.step(optimizer)` / `scaler.update()` (:800-801; with scaler=None a plain backward / step).  Returns the
    (device) loss tensor; no host sync inside."""
    if fused_prep:
        rgb, ir = prep_inputs(imgs_u8)
    else:  # the reference's three elementwise / copy passes
        f = imgs_u8.float() / 255.0
        rgb, ir = f[:, :3], f[:, 3:]
    with torch.autocast("cuda", dtype=autocast_dtype, enabled=autocast_dtype is not None):
        pred, comb = model(rgb, ir)
        loss, _ = compute_loss(pred, targets, comb.reshape(-1))  # a 0-d Combine_loss breaks len()
        if world_size > 1:
            loss = loss * world_size  # train.py:790-791
    if scaler is not None:
        scaler.scale(loss).backward()
        scaler.step(optimizer)
        scaler.update()
    else:
        loss.backward()
        optimizer.step()
    optimizer.zero_grad(set_to_none=True)
    return loss.detach()


@torch.no_grad()
def infer(model, rgb, ir, conf_thres: float = 0.25, iou_thres: float = 0.45, fused_post: bool = True):
    """detect_twostream.py:88-94: forward + NMS.  fused_post=True runs our batched post-processing (one NMS call for the whole
    batch instead of the per-image Python loop of utils/general.py:486-580; same detections)."""
    ref = import_reference()
    pred = model(rgb, ir)[0][0]
    if fused_post:
        from . import postprocess
        return postprocess.non_max_suppression(pred, conf_thres, iou_thres)
    return ref.general.non_max_suppression(pred, conf_thres, iou_thres)
