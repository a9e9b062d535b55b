#!/usr/bin/env python
"""torch.profiler breakdown of one detector training step (ours arm): where the step time goes.
Usage: python scripts/detector_profile.py [--size l --batch 16 --imgsz 640 --arm ours]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="l")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--imgsz", type=int, default=640)
ap.add_argument("--arm", default="ours")
ap.add_argument("--mode", default="train")
ap.add_argument("--channels-last", dest="cl", action="store_true")
ap.add_argument("--rows", type=int, default=45)
a = ap.parse_args()
from mmidet_b200 import harness as H  # noqa: E402

model = H.build_detector(a.size, a.arm, seed=0, channels_last=a.cl and a.mode == "train")
ref = H.import_reference()
imgs, targets = H.synthetic_batch(a.batch, a.imgsz, seed=1)
if a.mode == "train":
    model.train()
    hyp = H.scale_hyp(model, 6, a.imgsz)
    cl = ref.loss.ComputeLoss(model)
    opt = H.make_optimizer(model, hyp, a.batch)
    scaler = H.make_scaler(torch.float16)
    step = lambda: H.train_step(model, cl, opt, imgs, targets, autocast_dtype=torch.float16, fused_prep=a.arm == "ours", scaler=scaler)
else:
    model = H.prepare_inference(model, torch.float16, channels_last=a.cl)
    from mmidet_b200 import postprocess
    postprocess.install_detect(ref.yolo_test)

    def step():
        with torch.no_grad():
            rgb, ir = H.prep_inputs(imgs, torch.float16)
            return H.infer(model, rgb, ir, conf_thres=0.001)
for _ in range(3):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(3):
    step()
torch.cuda.synchronize()
print(f"wall per step: {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=a.rows, max_name_column_width=70))
