#!/usr/bin/env python
"""bench.py -- contract benchmark of the fusion hot path (fused selective scan forward + backward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one forward + one backward of the fused selective scan (+ SiLU gate) over one batch of synthetic
VIS+IR token features at the workload named in `config.workload` (BASELINE.json configs[2] at the shape the
north_star quotes: L=6400, d_inner=512; batch 16 per GPU = configs[3]'s per-GPU batch).  configs[1]/[3]/[4]
need the reference detector, which does not travel to the GPU box; they are parity-test territory.

  value     whole-job algorithmic GB/s (SURVEY 8d bytes formula x ranks / max-over-ranks device time), inputs
            resident in HBM, CUDA-event timed, L2 flushed between steps.
  e2e       the same metric through the host-buffer C-ABI entry (mmi_selscan_fwd_bwd_host): pinned host inputs,
            H2D + kernels + D2H inside the timed region.
  roofline  dominant kernel (the backward scan): algorithmic bytes / its event-timed duration vs measured HBM peak.
  cpu_baseline / --impl reference: the UNMODIFIED reference's own CPU path (MambaBlock.selective_scan + gate + autograd,
            imported from the staged checkout baseline/_ref) on a bounded sample of the same workload, all host cores
            (kind "reference"); where the checkout is not staged, the C + OpenMP oracle port (kind "port"), which is also
            reported as `cpu_port` for context (it is ~5x faster than the reference's PyTorch path on 16 cores).
  parity_relerr   batch entry 0 of the timed step's outputs against the fp64 oracle (the checker, after the timed region).
  general_A / bf16   the same step with a trained (non-geometric) A, and with bf16 I/O: what a training run sees after the
            first optimizer step / under autocast (secondary legs, a few steps each).
  train     BASELINE configs[3]: the data-parallel training step of the UNMODIFIED reference detector (two-stream
            YOLOv5l, 16 synthetic pairs per GPU, fp16 autocast + GradScaler as in train.py, ComputeLoss, SGD) with the CUDA fusion path plugged in,
            DDP gradient all-reduce over NCCL overlapped with the backward (8 MB buckets); `exposed_allreduce_ms` is the
            step time minus the same step under no_sync().  Needs the staged reference (baseline/_ref).
Only this file's cpu legs and tests/ touch oracle/; the product path is the CUDA library and fails loudly
without it."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(B=16, L=6400, ED=512, N=16, dtype="f32", gate=True)
METRIC = "fusion_scan_fwd_bwd_algorithmic_GBps"
UNIT = "GB/s"


def alg_bytes(B, L, ED, N, s, gate=True):
    """algorithmic bytes of the fused forward / backward (checkpoints and recompute excluded)."""
    fwd = B * L * ED * s * (4 if gate else 3) + B * L * N * s * 2 + (ED * N + ED) * 4
    bwd = B * L * ED * s * (7 if gate else 5) + B * L * N * s * 4 + 2 * (ED * N + ED) * 4
    return fwd, bwd


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------
def cpu_sample(nthreads=None, seconds_hint=12.0):
    """Oracle port (C, OpenMP) forward + backward on a bounded sample (B=1, full L and ED). Returns dict."""
    import numpy as np
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)  # torchrun exports OMP_NUM_THREADS=1 to its workers
    L, ED, N = WORKLOAD["L"], WORKLOAD["ED"], WORKLOAD["N"]
    Bs = 1
    rng = np.random.default_rng(0)
    x = rng.standard_normal((Bs, L, ED)).astype(np.float32)
    delta = np.log1p(np.exp(rng.standard_normal((Bs, L, ED)) - 3)).astype(np.float32)
    z = rng.standard_normal((Bs, L, ED)).astype(np.float32)
    Bm, Cm = rng.standard_normal((2, Bs, L, N)).astype(np.float32)
    dout = rng.standard_normal((Bs, L, ED)).astype(np.float32)
    A = -np.tile(np.arange(1, N + 1, dtype=np.float32), (ED, 1))
    D = np.ones(ED, np.float32)
    ol = O.lib()
    ol.oracle_set_threads(cores)  # the env var is only read when libgomp initialises; set it explicitly as well
    cores = int(ol.oracle_max_threads())
    fb, bb = alg_bytes(Bs, L, ED, N, 4)

    def once():
        t0 = time.perf_counter()
        O.selective_scan_fwd(x, delta, A, Bm, Cm, D, z=z)
        O.selective_scan_bwd(x, delta, A, Bm, Cm, D, dout, z=z)
        return time.perf_counter() - t0

    once()
    t = once()
    reps = max(1, min(8, int(seconds_hint / max(t, 1e-3)) - 1))
    ts = [t] + [once() for _ in range(reps)]
    best = min(ts)
    return {"value": round((fb + bb) / best / 1e9, 4), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle C/OpenMP port, fwd+bwd, B={Bs} L={L} ED={ED} N={N} fp32, best of {len(ts)}",
            "sec_per_sample": round(best, 4)}


REF_STAGED = os.path.join(ROOT, "baseline", "_ref")  # byte-for-byte copy of the reference (scripts/stage_reference.py)


class ReferenceCPU:
    """The UNMODIFIED reference's own CPU implementation of the path: MambaBlock.selective_scan (models/mamba.py:212-233,
    through models/pscan.py) + the SiLU gate + autograd backward, imported from the staged checkout baseline/_ref, on all
    host threads, on a bounded sample (B=1 of the workload's 16).  Raises if the checkout is not staged."""

    def __init__(self):
        import torch
        import torch.nn.functional as F
        if not os.path.isfile(os.path.join(REF_STAGED, "models", "mamba.py")):
            raise RuntimeError("baseline/_ref is not staged")
        if REF_STAGED not in sys.path:
            sys.path.insert(0, REF_STAGED)
        from models.mamba import MambaBlock, MambaConfig  # needs only torch
        self.torch, self.F = torch, F
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)  # torchrun exports OMP_NUM_THREADS=1 to its workers
        L, ED, N = WORKLOAD["L"], WORKLOAD["ED"], WORKLOAD["N"]
        g = torch.Generator().manual_seed(0)
        rn = lambda *sh: torch.randn(*sh, generator=g)
        x, z, self.dout = rn(1, L, ED), rn(1, L, ED), rn(1, L, ED)
        delta = F.softplus(rn(1, L, ED) - 3.0)
        Bm, Cm = rn(1, L, N), rn(1, L, N)
        A = -torch.arange(1, N + 1, dtype=torch.float32).repeat(ED, 1)
        self.leaves = [t.clone().requires_grad_(True) for t in (x, delta, z, A, Bm, Cm, torch.ones(ED))]
        self.blk = MambaBlock(MambaConfig(d_model=ED // 2, n_layers=1))
        self.sample = (f"unmodified reference (baseline/_ref): MambaBlock.selective_scan + gate + autograd backward on CPU, "
                       f"B=1 L={L} ED={ED} N={N} fp32, {torch.get_num_threads()} threads")

    def once(self):
        x, delta, z, A, Bm, Cm, D = self.leaves
        t0 = time.perf_counter()
        y = self.blk.selective_scan(x, delta, A, Bm, Cm, D)
        (y * self.F.silu(z)).backward(self.dout)
        dt = time.perf_counter() - t0
        for t in self.leaves:
            t.grad = None
        return dt


def reference_sample(n=3):
    """cpu_baseline of the GPU arm: the reference's own CPU path when staged, timed on a few samples."""
    r = ReferenceCPU()
    r.once()
    ts = [r.once() for _ in range(n)]
    fb, bb = alg_bytes(1, WORKLOAD["L"], WORKLOAD["ED"], WORKLOAD["N"], 4)
    best = min(ts)
    return {"value": round((fb + bb) / best / 1e9, 4), "unit": UNIT, "cores": r.cores, "kind": "reference",
            "sample": r.sample + f", best of {n}", "sec_per_sample": round(best, 4)}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation (staged checkout) on the host cores, rank 0 only; the C /
    OpenMP oracle port stands in only where the checkout is not staged."""
    if rank != 0:
        return
    steps, warm = args.steps, args.warmup
    ts = []
    try:
        r = ReferenceCPU()
        for i in range(warm + steps):
            t = r.once()
            if i >= warm:
                ts.append(t)
        base = {"cores": r.cores, "kind": "reference", "sample": r.sample}
    except Exception as e:
        base = None
        for i in range(warm + steps):
            q = cpu_sample(seconds_hint=0.0)
            if i >= warm:
                ts.append(q["sec_per_sample"])
            base = {"cores": q["cores"], "kind": "port", "sample": q["sample"] + f" (reference not staged: {e!r})"}
    fb, bb = alg_bytes(1, WORKLOAD["L"], WORKLOAD["ED"], WORKLOAD["N"], 4)
    mean_t = sum(ts) / len(ts)
    val = (fb + bb) / mean_t / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": round(mean_t * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(sample="each step = one bounded sample: B=1 of the workload's 16"),
            "cpu_baseline": {"value": round(val, 4), "unit": UNIT, "cores": base["cores"], "kind": base["kind"],
                             "sample": base["sample"]},
            "e2e": {"value": round(val, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(**extra):
    w = WORKLOAD
    c = {"workload": f"fused selective scan + SiLU gate, fwd+bwd, B={w['B']}/GPU L={w['L']} (80x80 tokens) d_inner={w['ED']} "
                     f"d_state={w['N']} fp32 I/O, A = S4D-real init of MambaBlock (models/mamba.py:158-159), "
                     f"delta=softplus(N(0,1)-3), x,z,B,C,dout~N(0,1) (BASELINE configs[2] at the north_star shape)",
         "B_per_gpu": w["B"], "L": w["L"], "d_inner": w["ED"], "d_state": w["N"],
         "l2": "256 MiB buffer rewritten between timed steps (and 1.5 GB working set >> 126 MB L2)",
         "parallelism": "one process per GPU, batch-sharded replicas, no data-path collective"}
    c.update(extra)
    return c


def parity_check(t, out, grads, b=0):
    """batch entry b of one step's outputs vs the fp64 oracle (checker only; after the timed region)."""
    import numpy as np
    from oracle import oracle as O
    f64 = lambda v: v[b:b + 1].detach().double().cpu().numpy()
    A, D = t["A"].double().cpu().numpy(), t["D"].double().cpu().numpy()
    a = {k: f64(t[k]) for k in ("x", "delta", "z", "Bm", "Cm", "dout")}
    ref = O.selective_scan_bwd(a["x"], a["delta"], A, a["Bm"], a["Cm"], D, a["dout"], z=a["z"], dtype=np.float64)
    ref["out"] = O.selective_scan_fwd(a["x"], a["delta"], A, a["Bm"], a["Cm"], D, z=a["z"], dtype=np.float64)
    got = dict(out=out, dx=grads[0], ddelta=grads[1], dz=grads[2], dB=grads[4], dC=grads[5])
    rel = lambda u, v: float(np.max(np.abs(u - v)) / max(np.max(np.abs(v)), 1e-30))
    return {k: float(f"{rel(f64(v), ref[k]):.3g}") for k, v in got.items()}


def detector_train_leg(rank, world, local_rank, dist, steps=8, warmup=5):
    """BASELINE configs[3] on the unmodified reference detector with the CUDA fusion path (see module docstring)."""
    import torch
    from mmidet_b200 import harness as H
    ref = H.import_reference()
    B, imgsz = 16, 640
    H.training_backend_flags()  # cudnn.benchmark = True, as the reference's train.py:66 sets it
    model = H.build_detector("l", "ours", seed=0, channels_last=True).train()
    hyp = H.scale_hyp(model, 6, imgsz)
    compute_loss = ref.loss.ComputeLoss(model)
    opt = H.make_optimizer(model, hyp, B * world)
    nparam = sum(p.numel() for p in model.parameters() if p.requires_grad)
    net = model
    if world > 1:
        from torch.nn.parallel import DistributedDataParallel as DDP
        # 8 MB buckets: the all-reduce of the early (head) buckets hides under the rest of the backward
        net = DDP(model, device_ids=[local_rank], output_device=local_rank, bucket_cap_mb=8, gradient_as_bucket_view=True,
                  broadcast_buffers=False)
    imgs, targets = H.synthetic_batch(B, imgsz, seed=100 + rank)
    scaler = H.make_scaler(torch.float16)  # the reference's recipe: fp16 autocast + GradScaler (train.py:706, :784-801)

    def run(n, sync=True):
        # sync=False: the same step on the bare module (no DDP hooks, no all-reduce) -- what this rank would do alone.
        # (net.no_sync() is not that baseline: with gradient_as_bucket_view and zero_grad(set_to_none=True) it re-allocates
        # every gradient outside the buckets each step and measures 40 % slower than the synchronised step.)
        target = net if sync else model
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            loss = H.train_step(target, compute_loss, opt, imgs, targets, autocast_dtype=torch.float16, world_size=world, scaler=scaler)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(loss)

    run(warmup)
    ms, loss = run(steps)
    leg = {"config": "two-stream YOLOv5l (unmodified reference Model / ComputeLoss, fusion = MambaFusion on the sm_100a kernels), "
                     "640x640 synthetic pairs, 16 / GPU, fp16 autocast + GradScaler (train.py:784-801), SGD, channels_last backbone",
           "pairs_per_s": round(world * B / (ms * 1e-3), 2), "ms_per_step": round(ms, 3), "n_gpus": world, "steps": steps,
           "params": nparam, "allreduce_bytes_per_step": nparam * 4 if world > 1 else 0, "loss": round(loss, 5)}
    if world > 1:
        run(2, sync=False)
        ms_nosync, _ = run(steps, sync=False)
        leg["ms_per_step_no_allreduce"] = round(ms_nosync, 3)
        leg["exposed_allreduce_ms"] = round(max(0.0, ms - ms_nosync), 3)  # (step-to-step noise is about +-2 ms)
    return leg


# ------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    from mmidet_b200 import _lib, ops
    lib = _lib.load()

    w = WORKLOAD
    B, L, ED, N = w["B"], w["L"], w["ED"], w["N"]
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    x, z, dout = rn(B, L, ED), rn(B, L, ED), rn(B, L, ED)
    delta = torch.nn.functional.softplus(rn(B, L, ED) - 3.0)
    Bm, Cm = rn(B, L, N), rn(B, L, N)
    A = -torch.arange(1, N + 1, device=dev, dtype=torch.float32).repeat(ED, 1)  # -exp(A_log), mamba.py:158-159,196
    D = torch.ones(ED, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    fb, bb = alg_bytes(B, L, ED, N, 4)

    hold = {}  # results of the newest step stay referenced until the next one replaces them -- in warm-up exactly as in the
    #            timed loop, so the caching allocator reaches its steady state before timing starts

    def step(ev=None):
        flush.zero_()  # L2 flush between steps, outside the event pairs
        if ev:
            ev[0].record()
        hold["out"], _, hold["chk"], hold["saved"] = ops.selscan_fwd_raw(x, delta, A, Bm, Cm, D, z=z, want_chk=True)
        if ev:
            ev[1].record()
        hold["grads"] = ops.selscan_bwd_raw(hold["saved"], hold["chk"], dout)
        if ev:
            ev[2].record()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()  # started before the warm-up: spawning nvidia-smi must not fall into the timed region
    for _ in range(max(args.warmup, 3)):
        step()
    # settle: a few more untimed steps until consecutive device times agree (clock ramp, allocator, tensor-map cache)
    prev = None
    for _ in range(30):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        step(e)
        torch.cuda.synchronize()
        cur = e[0].elapsed_time(e[2])
        if prev is not None and abs(cur - prev) <= 0.02 * prev:
            break
        prev = cur
    barrier()

    # ---- timed region: K steps, device-timed per step with the L2 flush outside the event pairs ----------------
    n0 = ops.launches
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        step(ev[k])
    barrier()
    out, grads = hold["out"], hold["grads"]
    launches = ops.launches - n0
    clocks = sampler.stop()
    t_f = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    t_b = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    t_step = t_f + t_b
    checksum = float(out.float().abs().mean()) + float(grads[0].float().abs().mean())

    # ---- secondary legs (untimed by the contract; a few steps each) -------------------------------------------------
    parity = None
    if rank == 0 and not args.no_cpu:
        try:
            parity = parity_check(dict(x=x, delta=delta, z=z, Bm=Bm, Cm=Cm, dout=dout, A=A, D=D), out, grads)
        except Exception as e:  # the oracle is a checker; its absence must not hide the GPU number
            parity = {"error": repr(e)}

    def quick(xx, dd, zz, BB, CC, gg, AA, n=5):
        def one():
            flush.zero_()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            e[0].record()
            o, _, ck, sv = ops.selscan_fwd_raw(xx, dd, AA, BB, CC, D, z=zz, want_chk=True)
            ops.selscan_bwd_raw(sv, ck, gg)
            e[1].record()
            torch.cuda.synchronize()
            return e[0].elapsed_time(e[1])
        for _ in range(3):
            one()
        ts = sorted(one() for _ in range(n))
        return ts[len(ts) // 2]

    A_tr = -torch.exp(torch.randn(ED, N, device=dev, generator=g) * 0.5 + 0.3)  # a trained A_log: no geometric rows
    t_gen = quick(x, delta, z, Bm, Cm, dout, A_tr)
    hb = lambda t: t.to(torch.bfloat16)
    t_bf = quick(hb(x), hb(delta), hb(z), hb(Bm), hb(Cm), hb(dout), A)
    fb16, bb16 = alg_bytes(B, L, ED, N, 2)

    # ---- end-to-end through the host-buffer C ABI (pinned host memory, H2D + kernels + D2H) -------------------
    e2e = None
    e2e_steps = max(2, min(args.steps, 5))
    if not args.no_e2e:
        import ctypes
        pin = lambda t: t.detach().cpu().pin_memory()
        hx, hd, hz, hB, hC, hg, hA, hD = (pin(t) for t in (x, delta, z, Bm, Cm, dout, A, D))
        ho, hdx, hdd, hdz = (torch.empty_like(hx).pin_memory() for _ in range(4))
        hdB, hdC = torch.empty_like(hB).pin_memory(), torch.empty_like(hC).pin_memory()
        hdA, hdD = torch.empty_like(hA).pin_memory(), torch.empty_like(hD).pin_memory()
        P = lambda t: ctypes.c_void_p(t.data_ptr())

        def host_step():
            _lib.check(lib.mmi_selscan_fwd_bwd_host(P(hx), P(hd), P(hz), P(hA), P(hB), P(hC), P(hD), P(hg), P(ho), P(hdx),
                                                    P(hdd), P(hdz), P(hdA), P(hdB), P(hdC), P(hdD), B, L, ED, N,
                                                    _lib.MMI_F32, 0), "mmi_selscan_fwd_bwd_host")

        host_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_step()  # synchronises its own stream before returning
        barrier()
        t_e2e = (time.perf_counter() - t0) / e2e_steps
        h2d = sum(t.numel() * t.element_size() for t in (hx, hd, hz, hB, hC, hg, hA, hD))
        d2h = sum(t.numel() * t.element_size() for t in (ho, hdx, hdd, hdz, hdB, hdC, hdA, hdD))
        e2e_check = float(ho.abs().mean()) + float(hdx.abs().mean())
        lib.mmi_host_workspace_free()
        e2e = (t_e2e, h2d, d2h, e2e_check)

    # ---- max over ranks -------------------------------------------------------------------------------------
    times = torch.tensor([t_step, t_f, t_b, e2e[0] * 1e3 if e2e else 0.0, t_gen, t_bf], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_step, t_f, t_b, t_e2e_ms, t_gen, t_bf = [float(v) for v in times]

    train_leg = None
    if not args.no_train:
        hold.clear()
        del out, grads
        torch.cuda.empty_cache()
        try:
            train_leg = detector_train_leg(rank, world, local_rank, dist)
        except Exception as e:  # e.g. the reference checkout was not staged: the scan numbers still stand
            train_leg = {"error": repr(e)[:300]}

    if rank == 0:
        peak, peak_src = measured_peak()
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("selscan_bwd2_kernel", {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        val = world * (fb + bb) / (t_step * 1e-3) / 1e9
        line = {"metric": METRIC, "value": round(val, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": round(t_step, 4), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(),
                "frac_of_hbm_peak": round(val / world / peak, 4),
                "roofline": {"bound": "hbm", "kernel": "selscan_bwd2_kernel<float, true> (+ partial-reduction kernel)",
                             "achieved": round(bb / (t_b * 1e-3) / 1e9, 2), "peak": peak, "unit": "GB/s",
                             "frac": round(bb / (t_b * 1e-3) / 1e9 / peak, 4), "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": bb, "ms_per_launch": round(t_b, 4)},
                "roofline_fwd": {"bound": "hbm", "kernel": "selscan_fwd_kernel<float>",
                                 "achieved": round(fb / (t_f * 1e-3) / 1e9, 2), "peak": peak, "unit": "GB/s",
                                 "frac": round(fb / (t_f * 1e-3) / 1e9 / peak, 4), "algorithmic_bytes_per_launch": fb,
                                 "ms_per_launch": round(t_f, 4)},
                "general_A": {"what": "same step, trained (non-geometric) A: 16 exponentials per step instead of 1",
                              "value": round(world * (fb + bb) / (t_gen * 1e-3) / 1e9, 2), "unit": UNIT, "ms_per_step": round(t_gen, 4),
                              "frac_of_hbm_peak": round((fb + bb) / (t_gen * 1e-3) / 1e9 / peak, 4)},
                "roofline_bf16": {"what": "same step, bf16 I/O (fp32 state): half the algorithmic bytes, same arithmetic",
                                  "achieved": round((fb16 + bb16) / (t_bf * 1e-3) / 1e9, 2), "peak": peak, "unit": "GB/s",
                                  "frac": round((fb16 + bb16) / (t_bf * 1e-3) / 1e9 / peak, 4), "ms_per_step": round(t_bf, 4)},
                "parity_relerr": parity, "train": train_leg,
                "clocks": clocks, "gpu_launches": launches, "checksum": round(checksum, 6)}
        if e2e:
            line["e2e"] = {"value": round(world * (fb + bb) / (t_e2e_ms * 1e-3) / 1e9, 3), "unit": UNIT,
                           "h2d_bytes_per_step": e2e[1], "d2h_bytes_per_step": e2e[2], "ms_per_step": round(t_e2e_ms, 3),
                           "steps": e2e_steps, "api": "mmi_selscan_fwd_bwd_host (C ABI, pinned host buffers)",
                           "checksum": round(e2e[3], 6)}
        if world == 1 and not args.no_cpu:
            try:
                line["cpu_port"] = cpu_sample(seconds_hint=4.0)  # the C / OpenMP restatement, for context
            except Exception as e:  # the oracle is a checker; its absence must not hide the GPU number
                line["cpu_port"] = {"error": repr(e)}
            try:
                line["cpu_baseline"] = reference_sample()  # the reference's own PyTorch CPU path (staged checkout)
            except Exception:
                line["cpu_baseline"] = line["cpu_port"]
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the detector DDP training leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517"), __file__,
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup)]
        cmd += ["--no-e2e"] if args.no_e2e else []
        cmd += ["--no-train"] if args.no_train else []
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
