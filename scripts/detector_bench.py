#!/usr/bin/env python
"""scripts/detector_bench.py -- BASELINE configs[1] / [3] / [4] on the UNMODIFIED reference detector (harness.py).

    python scripts/detector_bench.py logits [--size s] [--imgsz 640]
        two-stream YOLOv5s forward, batch 1: Detect output of the CUDA fusion path vs the same model on the reference's
        pure-PyTorch MambaBlock / pscan, both on the GPU, same weights -> rel-err + latency of both arms        (configs[1])
    python [-m torch.distributed.run ...] scripts/detector_bench.py train [--size l] [--batch 16] [--arm ours|pytorch]
        training step (uint8 batch -> /255 -> split, fp16 autocast forward + GradScaler as train.py:784-801, ComputeLoss, backward, DDP gradient
        all-reduce, SGD step), 16 pairs / GPU, torchrun at 1/2/4/8 -> image-pairs / s                              (configs[3])
    python scripts/detector_bench.py infer [--size x] [--imgsz 1280] [--batch 32]
        detect_twostream.py's timing window (model forward + NMS), fp16 -> pairs / s, p50 / p99 latency           (configs[4])
Each mode prints one JSON line (rank 0)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def timed(fn, warm, steps):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts


def mode_logits(args):
    from mmidet_b200 import harness as H
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref_model = H.build_detector(args.size, "pytorch", seed=0).eval()
    our_model = H.build_detector(args.size, "ours", seed=0, state_dict=ref_model.state_dict()).eval()
    g = torch.Generator().manual_seed(2)
    rgb = torch.rand(args.batch, 3, args.imgsz, args.imgsz, generator=g).cuda()
    ir = torch.rand(args.batch, 3, args.imgsz, args.imgsz, generator=g).cuda()
    with torch.no_grad():
        (zr, xr), _ = ref_model(rgb, ir)
        (zo, xo), _ = our_model(rgb, ir)
        t_ref = timed(lambda: ref_model(rgb, ir), 3, args.steps)
        t_our = timed(lambda: our_model(rgb, ir), 3, args.steps)
    line = {"mode": "logits", "config": f"two-stream YOLOv5{args.size} forward, {args.imgsz}x{args.imgsz} synthetic pair, batch {args.batch}, fp32",
            "detect_shape": list(zo.shape), "relerr_detect": relerr(zo, zr),
            "relerr_raw_maps": [relerr(a, b) for a, b in zip(xo, xr)],
            "ms_ours_p50": t_our[len(t_our) // 2], "ms_pytorch_gpu_p50": t_ref[len(t_ref) // 2],
            "speedup_vs_pytorch_gpu": t_ref[len(t_ref) // 2] / t_our[len(t_our) // 2]}
    print(json.dumps(line), flush=True)


def mode_train(args):
    from mmidet_b200 import harness as H
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    lr = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    if not args.no_cudnn_benchmark:
        H.training_backend_flags()  # cudnn.benchmark = True, as the reference's train.py sets it
    model = H.build_detector(args.size, args.arm, seed=0, channels_last=args.channels_last).train()
    hyp = H.scale_hyp(model, 6, args.imgsz)
    ref = H.import_reference()
    compute_loss = ref.loss.ComputeLoss(model)
    opt = H.make_optimizer(model, hyp, args.batch * world)
    nparam = sum(p.numel() for p in model.parameters() if p.requires_grad)
    net = model
    if world > 1:
        from torch.nn.parallel import DistributedDataParallel as DDP
        net = DDP(model, device_ids=[lr], output_device=lr, bucket_cap_mb=args.bucket_mb, gradient_as_bucket_view=True,
                  broadcast_buffers=False, static_graph=True)
    imgs, targets = H.synthetic_batch(args.batch, args.imgsz, seed=100 + rank)
    ac = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": None}[args.autocast]

    scaler = H.make_scaler(ac)  # fp16: the reference's GradScaler (train.py:706, :796-801)

    def step():
        return H.train_step(net, compute_loss, opt, imgs, targets, autocast_dtype=ac, world_size=world,
                            fused_prep=args.arm == "ours", scaler=scaler)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        loss = step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t[0])
        print(json.dumps({"mode": "train", "arm": args.arm, "n_gpus": world,
                          "config": f"two-stream YOLOv5{args.size} training step, {args.imgsz}x{args.imgsz} synthetic pairs, batch "
                                    f"{args.batch}/GPU, autocast {args.autocast}{' + GradScaler' if scaler is not None else ''}, SGD, DDP bucket {args.bucket_mb} MB, "
                                    f"{'channels_last' if args.channels_last else 'NCHW'} backbone",
                          "pairs_per_s": round(world * args.batch / (ms * 1e-3), 2), "ms_per_step": round(ms, 3),
                          "params": nparam, "allreduce_bytes_per_step": nparam * 4 if world > 1 else 0,
                          "loss": float(loss), "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def mode_infer(args):
    from mmidet_b200 import harness as H
    from mmidet_b200 import postprocess
    ref = H.import_reference()
    dt = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[args.dtype]
    model = H.prepare_inference(H.build_detector(args.size, args.arm, seed=0), dt, fuse=not args.no_fuse,
                                channels_last=args.channels_last)
    if args.arm == "ours":
        postprocess.install_detect(ref.yolo_test)
    imgs, _ = H.synthetic_batch(args.batch, args.imgsz, seed=7)

    def once():
        if args.arm == "ours":
            rgb, ir = H.prep_inputs(imgs, dt)
        else:  # detect_twostream.py:74-85
            f = imgs.to(dt) / 255.0
            rgb, ir = f[:, :3], f[:, 3:]
        return H.infer(model, rgb, ir, conf_thres=args.conf, fused_post=args.arm == "ours")

    with torch.no_grad():
        det = once()
        ts = timed(once, args.warmup, args.steps)
    p50, p99 = ts[len(ts) // 2], ts[min(len(ts) - 1, int(len(ts) * 0.99))]
    print(json.dumps({"mode": "infer", "arm": args.arm,
                      "config": f"two-stream YOLOv5{args.size} inference (input prep + forward + NMS, detect_twostream.py:74-94), "
                                f"{args.imgsz}x{args.imgsz}, batch {args.batch}, {args.dtype}, Conv+BN {'fused' if not args.no_fuse else 'unfused'}, "
                                f"{'channels_last' if args.channels_last else 'NCHW'} backbone",
                      "pairs_per_s": round(args.batch / (p50 * 1e-3), 2), "ms_p50": round(p50, 3), "ms_p99": round(p99, 3),
                      "detections": int(sum(d.shape[0] for d in det)),
                      "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["logits", "train", "infer"])
    ap.add_argument("--size", default=None)
    ap.add_argument("--imgsz", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--arm", default="ours", choices=["ours", "pytorch"])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--autocast", default="fp16", choices=["fp16", "bf16", "fp32"],
                    help="fp16 (+ GradScaler) is the reference's own recipe, train.py:706/:784")
    ap.add_argument("--dtype", default="fp16")
    ap.add_argument("--conf", type=float, default=0.25)
    ap.add_argument("--bucket-mb", dest="bucket_mb", type=int, default=8)
    ap.add_argument("--channels-last", dest="channels_last", action="store_true")
    ap.add_argument("--no-fuse", dest="no_fuse", action="store_true")
    ap.add_argument("--no-cudnn-benchmark", dest="no_cudnn_benchmark", action="store_true")
    args = ap.parse_args()
    d = {"logits": ("s", 640, 1), "train": ("l", 640, 16), "infer": ("x", 1280, 32)}[args.mode]
    args.size = args.size or d[0]
    args.imgsz = args.imgsz or d[1]
    args.batch = args.batch or d[2]
    globals()["mode_" + args.mode](args)


if __name__ == "__main__":
    main()
