"""Detector input / post-processing on the kernels of csrc/detect.cu.

    split_normalize(imgs_u8)              train.py:743-745, detect_twostream.py:74-85   uint8 (B,6,H,W) -> rgb, ir float / 255
    detect_forward(self, x)               Detect.forward (models/yolo_test.py:47-68); bound by install_detect(): the training
                                          branch is the reference's own code path, the inference branch decodes every level
                                          with one kernel each straight into the concatenated prediction
    non_max_suppression(pred, conf, iou)  utils/general.py:486-580 (best-class branch: multi_label=False, no apriori labels,
                                          no merge) for the whole batch: candidates, one stable sort, suppression bit matrix,
                                          per-image sweep.  Returns the reference's list of (n_i, 6) tensors.
CUDA tensors only; no eager fallback."""
from __future__ import annotations

import torch

from . import _lib
from . import ops as _ops

MAX_WH, MAX_DET, MAX_NMS = 4096.0, 300, 30000  # utils/general.py:499-501


def split_normalize(imgs_u8: torch.Tensor, dtype=torch.float32):
    if not imgs_u8.is_cuda:
        raise RuntimeError("mmidet_b200.split_normalize: CUDA tensor required (no CPU path)")
    if imgs_u8.dtype != torch.uint8 or imgs_u8.dim() != 4 or imgs_u8.shape[1] != 6:
        raise ValueError(f"split_normalize: expected a uint8 (B, 6, H, W) batch, got {imgs_u8.dtype} {tuple(imgs_u8.shape)}")
    lib = _lib.load()
    src = imgs_u8.contiguous()
    B, _, H, W = src.shape
    rgb = torch.empty((B, 3, H, W), dtype=dtype, device=src.device)
    ir = torch.empty((B, 3, H, W), dtype=dtype, device=src.device)
    _lib.check(lib.mmi_u8_split_normalize(_ops._ptr(src), _ops._ptr(rgb), _ops._ptr(ir), B, H, W, _ops._DT[dtype],
                                          _ops._stream(src)), "mmi_u8_split_normalize")
    _ops.launches += 1
    return rgb, ir


def detect_forward(self, x):
    """Drop-in for Detect.forward (models/yolo_test.py:47-68).  Training: the reference's view / permute / contiguous per
    level (autograd needs them).  Inference: (torch.cat(z, 1), x) with x[i] the permuted raw maps, as the reference."""
    self.training |= self.export
    if self.training:
        for i in range(self.nl):
            x[i] = self.m[i](x[i])
            bs, _, ny, nx = x[i].shape
            x[i] = x[i].view(bs, self.na, self.no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
        return x
    lib = _lib.load()
    maps = [self.m[i](x[i]).contiguous() for i in range(self.nl)]
    bs, dt, dev = maps[0].shape[0], maps[0].dtype, maps[0].device
    if not maps[0].is_cuda or dt not in _ops._DT:
        raise RuntimeError("mmidet_b200.detect_forward: CUDA float tensors required (no CPU path)")
    rows = [self.na * m.shape[2] * m.shape[3] for m in maps]
    pred = torch.empty((bs, sum(rows), self.no), dtype=dt, device=dev)
    off = 0
    for i, m in enumerate(maps):
        ny, nx = m.shape[2], m.shape[3]
        raw = torch.empty((bs, self.na, ny, nx, self.no), dtype=dt, device=dev)
        anchor_wh = self.anchor_grid[i].reshape(self.na, 2).float().contiguous()
        _lib.check(lib.mmi_detect_decode(_ops._ptr(m), _ops._ptr(raw), _ops._ptr(pred), bs, self.na, self.no, ny, nx,
                                         float(self.stride[i]), _ops._ptr(anchor_wh), sum(rows), off, _ops._DT[dt],
                                         _ops._stream(m)), "mmi_detect_decode")
        _ops.launches += 1
        x[i] = raw
        off += rows[i]
    return pred, x


def install_detect(ref_yolo_test):
    """Bind detect_forward onto the reference's Detect class (class-attribute patch, state_dict untouched).  Returns the
    (obj, attr, old) triple list that mamba.uninstall() accepts."""
    old = ref_yolo_test.Detect.forward
    ref_yolo_test.Detect.forward = detect_forward
    return [(ref_yolo_test.Detect, "forward", old)]


@torch.no_grad()
def non_max_suppression(prediction: torch.Tensor, conf_thres: float = 0.25, iou_thres: float = 0.45, max_det: int = MAX_DET):
    """prediction (bs, rows, 5 + nc) as Detect returns it -> list of bs tensors (n_i, 6) = x1, y1, x2, y2, conf, cls, the
    kept detections of utils/general.py:486-580 in the same order (descending confidence)."""
    if not prediction.is_cuda:
        raise RuntimeError("mmidet_b200.non_max_suppression: CUDA tensor required (no CPU path)")
    if prediction.dim() != 3 or prediction.shape[2] < 6 or prediction.dtype not in _ops._DT:
        raise ValueError(f"non_max_suppression: expected a float (bs, rows, 5 + nc) prediction, got {tuple(prediction.shape)}")
    lib = _lib.load()
    pred = prediction.contiguous()
    bs, rows, no = pred.shape
    dev = pred.device
    det = torch.empty((bs * rows, 6), dtype=torch.float32, device=dev)
    key = torch.empty(bs * rows, dtype=torch.float64, device=dev)
    count = torch.empty(bs, dtype=torch.int32, device=dev)
    st = _ops._stream(pred)
    _lib.check(lib.mmi_nms_candidates(_ops._ptr(pred), _ops._ptr(det), _ops._ptr(key), _ops._ptr(count), bs, rows, no,
                                      float(conf_thres), _ops._DT[pred.dtype], st), "mmi_nms_candidates")
    order = torch.sort(key, stable=True).indices  # image-major, descending confidence, ties in row order (as the reference)
    counts = count.tolist()  # the one host sync of the post-processing (the reference syncs once per image)
    max_count = max(counts)
    start = torch.zeros(bs, dtype=torch.int32, device=dev)
    if bs > 1:
        start[1:] = torch.cumsum(count[:-1], 0)
    words = (min(max_count, MAX_NMS) + 63) // 64
    total = sum(counts)
    if total * words * 8 > (8 << 30):
        raise RuntimeError(f"non_max_suppression: {total} candidates need a {total * words * 8 >> 20} MiB suppression matrix; "
                           "raise conf_thres")
    mask = torch.empty(max(total * words, 1), dtype=torch.int64, device=dev)
    keep = torch.empty((bs, max_det), dtype=torch.int64, device=dev)
    nkeep = torch.empty(bs, dtype=torch.int32, device=dev)
    _lib.check(lib.mmi_nms_suppress(_ops._ptr(det), _ops._ptr(order), _ops._ptr(start), _ops._ptr(count), _ops._ptr(mask),
                                    _ops._ptr(keep), _ops._ptr(nkeep), bs, max_count, MAX_NMS, max_det, float(iou_thres),
                                    MAX_WH, st), "mmi_nms_suppress")
    _ops.launches += 3
    nk = nkeep.tolist()
    return [det[keep[i, :nk[i]]] for i in range(bs)]
