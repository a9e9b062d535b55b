"""GPU parity of the detector input / post-processing kernels (csrc/detect.cu) against torch
restatements of the reference lines they replace, and against the reference's own functions when its checkout is staged."""
import numpy as np
import pytest
import torch

from tests.util import relerr

pytestmark = pytest.mark.gpu


def _ref():
    from mmidet_b200 import harness
    try:
        return harness.import_reference()
    except RuntimeError:
        return None


@pytest.mark.parametrize("shape", [(2, 6, 64, 64), (1, 6, 33, 17), (3, 6, 160, 96)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_split_normalize(shape, dtype):
    """train.py:743-745: imgs.float() / 255 then the RGB | IR channel split -- bit-exact in fp32."""
    from mmidet_b200.postprocess import split_normalize
    g = torch.Generator().manual_seed(shape[2])
    imgs = torch.randint(0, 256, shape, dtype=torch.uint8, generator=g).cuda()
    rgb, ir = split_normalize(imgs, dtype)
    f = imgs.float() / 255.0
    want = imgs.to(dtype) / 255.0  # detect_twostream.py:78-79: the division happens on the already-cast tensor
    assert torch.equal(rgb, want[:, :3]) and torch.equal(ir, want[:, 3:])  # bit-exact in every dtype
    if dtype == torch.float32:
        assert torch.equal(rgb, f[:, :3]) and torch.equal(ir, f[:, 3:])
    assert rgb.is_contiguous() and ir.is_contiguous() and rgb.dtype == dtype


def _detect_reference(xs, anchors, strides, na, no):
    """models/yolo_test.py:47-68, inference branch, restated on already-convolved maps."""
    z, raws = [], []
    for i, x in enumerate(xs):
        bs, _, ny, nx = x.shape
        r = x.view(bs, na, no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
        yv, xv = torch.meshgrid([torch.arange(ny), torch.arange(nx)], indexing="ij")
        grid = torch.stack((xv, yv), 2).view(1, 1, ny, nx, 2).float().to(x.device)
        y = r.sigmoid()
        y[..., 0:2] = (y[..., 0:2] * 2. - 0.5 + grid) * strides[i]
        y[..., 2:4] = (y[..., 2:4] * 2) ** 2 * anchors[i].view(1, na, 1, 1, 2)
        z.append(y.view(bs, -1, no))
        raws.append(r)
    return torch.cat(z, 1), raws


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.float16, 2e-3)])
@pytest.mark.parametrize("bs,nc,sizes", [(2, 6, [(20, 20), (10, 10), (5, 5)]), (1, 80, [(7, 9), (4, 5), (2, 3)]), (3, 1, [(33, 31)])])
def test_detect_decode(bs, nc, sizes, dtype, tol):
    import ctypes
    from mmidet_b200 import _lib, ops
    lib = _lib.load()
    na, no = 3, nc + 5
    torch.manual_seed(nc)
    xs = [(torch.randn(bs, na * no, ny, nx, device="cuda") * 2).to(dtype) for ny, nx in sizes]
    anchors = [torch.rand(na, 2, device="cuda") * 100 + 5 for _ in sizes]
    strides = [8.0 * 2 ** i for i in range(len(sizes))]
    want, raws = _detect_reference(xs, anchors, strides, na, no)
    rows = [na * ny * nx for ny, nx in sizes]
    pred = torch.full((bs, sum(rows), no), float("nan"), device="cuda", dtype=dtype)
    off = 0
    for i, x in enumerate(xs):
        ny, nx = sizes[i]
        raw = torch.empty((bs, na, ny, nx, no), device="cuda", dtype=dtype)
        _lib.check(lib.mmi_detect_decode(ops._ptr(x), ops._ptr(raw), ops._ptr(pred), bs, na, no, ny, nx, strides[i],
                                         ops._ptr(anchors[i].contiguous()), sum(rows), off, ops._DT[dtype], ops._stream(x)), "decode")
        assert torch.equal(raw, raws[i])
        off += rows[i]
    assert relerr(pred.float().cpu().numpy(), want.float().cpu().numpy()) <= tol


def _nms_reference(prediction, conf_thres, iou_thres, max_det=300, max_wh=4096, max_nms=30000):
    """utils/general.py:486-580 (best-class branch) restated with torchvision.ops.nms, per image."""
    import torchvision
    out = []
    xc = prediction[..., 4] > conf_thres
    for xi, x in enumerate(prediction):
        x = x[xc[xi]].clone()
        if not x.shape[0]:
            out.append(torch.zeros((0, 6), device=prediction.device))
            continue
        x[:, 5:] *= x[:, 4:5]
        box = x[:, :4].clone()
        box[:, 0] = x[:, 0] - x[:, 2] / 2
        box[:, 1] = x[:, 1] - x[:, 3] / 2
        box[:, 2] = x[:, 0] + x[:, 2] / 2
        box[:, 3] = x[:, 1] + x[:, 3] / 2
        conf, j = x[:, 5:].max(1, keepdim=True)
        x = torch.cat((box, conf, j.float()), 1)[conf.view(-1) > conf_thres]
        if not x.shape[0]:
            out.append(torch.zeros((0, 6), device=prediction.device))
            continue
        if x.shape[0] > max_nms:
            x = x[x[:, 4].argsort(descending=True)[:max_nms]]
        c = x[:, 5:6] * max_wh
        i = torchvision.ops.nms(x[:, :4] + c, x[:, 4], iou_thres)[:max_det]
        out.append(x[i])
    return out


def _fake_prediction(bs, rows, nc, seed, dtype=torch.float32, dense=False):
    g = torch.Generator().manual_seed(seed)
    p = torch.empty(bs, rows, 5 + nc)
    centres = torch.rand(bs, 12, 2, generator=g) * 500 + 70  # a dozen object clusters per image -> many overlapping boxes
    pick = torch.randint(0, 12, (bs, rows), generator=g)
    p[..., 0:2] = torch.gather(centres, 1, pick[..., None].expand(-1, -1, 2)) + torch.randn(bs, rows, 2, generator=g) * 6
    p[..., 2:4] = torch.rand(bs, rows, 2, generator=g) * 80 + 20
    p[..., 4] = torch.rand(bs, rows, generator=g) ** (1 if dense else 4)
    p[..., 5:] = torch.rand(bs, rows, nc, generator=g)
    return p.to(dtype).cuda()


@pytest.mark.parametrize("case", [(2, 1575, 6, 0.25, False), (5, 25200, 6, 0.25, False), (3, 4000, 80, 0.001, True),
                                  (1, 300, 1, 0.25, False), (4, 2000, 6, 0.9999, False)])
def test_nms_matches_reference_restatement(case):
    """kept boxes, their order and their count per image == the per-image torchvision-based reference algorithm,
    including > 300 survivors (max_det cut), an image without candidates, and one class."""
    from mmidet_b200.postprocess import non_max_suppression
    bs, rows, nc, conf, dense = case
    pred = _fake_prediction(bs, rows, nc, seed=rows + nc, dense=dense)
    if bs > 2:
        pred[1, :, 4] = 0.0  # an image with no candidates at all
    got = non_max_suppression(pred, conf, 0.45)
    want = _nms_reference(pred, conf, 0.45)
    assert len(got) == len(want) == bs
    for a, b in zip(got, want):
        assert a.shape == b.shape, (a.shape, b.shape)
        assert torch.equal(a, b)


def test_nms_matches_the_reference_function():
    """against utils.general.non_max_suppression of the staged reference checkout itself (fp32 and fp16 predictions)."""
    ref = _ref()
    if ref is None:
        pytest.skip("reference checkout not staged (scripts/stage_reference.py)")
    from mmidet_b200.postprocess import non_max_suppression
    for dtype in (torch.float32, torch.float16):
        pred = _fake_prediction(4, 6300, 6, seed=3, dtype=dtype)
        got = non_max_suppression(pred.clone(), 0.25, 0.45)
        want = ref.general.non_max_suppression(pred.clone(), 0.25, 0.45)
        for a, b in zip(got, want):
            assert a.shape == b.shape and torch.equal(a, b.float())


def test_detect_forward_dropin_matches_reference_module():
    """Detect.forward of the staged reference vs our drop-in bound onto the same module instance (eval: decoded prediction
    and raw maps; train: the raw maps)."""
    ref = _ref()
    if ref is None:
        pytest.skip("reference checkout not staged (scripts/stage_reference.py)")
    from mmidet_b200 import mamba, postprocess
    Y = ref.yolo_test
    torch.manual_seed(0)
    anchors = [[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]]
    det = Y.Detect(nc=6, anchors=anchors, ch=(32, 64, 128)).cuda()
    det.stride = torch.tensor([8.0, 16.0, 32.0])
    feats = lambda: [torch.randn(2, c, s, s, device="cuda") for c, s in ((32, 20), (64, 10), (128, 5))]
    torch.manual_seed(1)
    f = feats()
    for training in (False, True):
        det.train(training)
        with torch.no_grad():
            want = det([t.clone() for t in f])
            saved = postprocess.install_detect(Y)
            try:
                got = det([t.clone() for t in f])
            finally:
                mamba.uninstall(saved)
        if training:
            for a, b in zip(got, want):
                assert torch.equal(a, b)
        else:
            assert relerr(got[0].cpu().numpy(), want[0].cpu().numpy()) <= 2e-6
            for a, b in zip(got[1], want[1]):
                assert torch.equal(a, b)
