#!/usr/bin/env python
"""Benchmark of the VSPW per-clip hot path: TCB-PSP ResNet101-dilated, T=5 synthetic 480p clips, train fwd+bwd.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16x3|bf16]

One JSON line on stdout (rank 0).  `value` = clip-frames/s with the clip already resident in HBM, `e2e` = the same
step driven through the reference-facing module call from pinned HOST buffers (H2D of images+labels and D2H of
the loss inside the timed region).  `roofline` = the dominant kernel family (implicit-GEMM convolutions) timed with
CUDA events on the launching stream inside the timed steps.  `cpu_baseline` / `--impl reference` time the CPU
oracle (oracle/tcb_oracle.py = the reference's PyTorch-CPU path restated) on the box's host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _oracle():
    """The CPU oracle (test infrastructure) — imported ONLY by the cpu_baseline / --impl reference / gpu_library_baseline
    legs, never by the product arm's timed path."""
    od = os.path.join(ROOT, "oracle")
    if od not in sys.path:
        sys.path.insert(0, od)
    import tcb_oracle
    return tcb_oracle

import torch  # noqa: E402

METRIC = "clip-frames/sec @480p T=5 ResNet101-TCB-PSP"
UNIT = "clip-frames/s"
T_FRAMES, N_CLIPS, H, W, NUM_CLASS = 5, 2, 480, 854, 124
# algorithmic FLOPs (2*MAC, conv+matmul) of one TCB-PSP train step at this config, BASELINE.md section 3
STEP_TFLOP = 20.631


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("VSPW_PRECISION", "bf16x3"), choices=["fp32", "bf16x3", "bf16"])
    ap.add_argument("--syncbn", default="auto", choices=["auto", "on", "off"],
                    help="cross-rank BN statistics (the reference's multi-GPU semantics and train_clip2.py's default); auto = on when N > 1")
    ap.add_argument("--syncbn-exchange", default="peer", choices=["peer", "nccl"],
                    help="peer = one-shot exchange over NVLink peer memory per BN layer (csrc/peer.cu); nccl = a library all-reduce per layer")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-library-baseline", action="store_true", help="skip the stated-context leg: the oracle port on this GPU through cuDNN")
    ap.add_argument("--profile-run", action="store_true", help="for ncu captures only: honour --warmup below 3, skip e2e/cpu legs")
    ap.add_argument("--cpu-sample", default="240x427", help="HxW of the bounded CPU sample")
    ap.add_argument("--model", default="psp", choices=["psp", "ocr"], help="psp = TCB-PSP (the headline metric, BASELINE configs[1]); ocr = TCB-OCR (configs[2], reported under its own metric name)")
    ap.add_argument("--torch-sgd", action="store_true", help="use torch.optim.SGD instead of the fused vspw_sgd_momentum_step (same update rule)")
    ap.add_argument("--size", default="", help="HxW override, only with --profile-run (host-overhead probes); never a bench value")
    ap.add_argument("--kernel-profile", default="", help="write a per-entry-point CUDA-event breakdown of 2 extra steps to this file")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# --------------------------------------------------------------------------------------------------
def build_model(device, seed=0, kind="psp"):
    from cvpr2021_vspw_implement_b200 import models as M
    torch.manual_seed(seed)
    ns = argparse.Namespace(num_class=NUM_CLASS, psp_weight=False, use_memory=False, memory_num=8, clipocr_all=False)
    enc = M.ModelBuilder.build_encoder("resnet101dilated")
    cls = M.Clip_PSP if kind == "psp" else M.ClipOCRNet
    m = cls(enc, torch.nn.NLLLoss(ignore_index=255), ns, deep_sup_scale=0.4)
    return m.to(device).train()


def make_optimizer(m, lr=0.002, fused=True):
    """create_optimizers of the reference (train_clip2.py:215-236): SGD momentum .9, 4 param groups.  The duplicate
    yields of the generators (quirk Q10) are de-duplicated here because torch >= 2 rejects duplicate parameters."""
    def uniq(gen, seen):
        out = []
        for p in gen:
            if id(p) not in seen:
                seen.add(id(p))
                out.append(p)
        return out
    seen = set()
    groups = [
        {"params": uniq(m.get_1x_lr_params(), seen), "lr": lr * 0.1, "weight_decay": 1e-4},
        {"params": uniq(m.get_10x_lr_params(), seen), "lr": lr, "weight_decay": 1e-4},
        {"params": uniq(m.get_1x_lr_params_bias(), seen), "lr": lr * 0.1, "weight_decay": 0.0},
        {"params": uniq(m.get_10x_lr_params_bias(), seen), "lr": lr, "weight_decay": 0.0},
    ]
    if fused:
        from cvpr2021_vspw_implement_b200.optim import FusedSGD
        return FusedSGD([g for g in groups if g["params"]], lr=lr, momentum=0.9)
    return torch.optim.SGD([g for g in groups if g["params"]], lr=lr, momentum=0.9)


def feed_from(imgs, labs):
    return {"img_data": imgs[0], "seg_label": labs[0], "clipimgs_data": list(imgs[1:]), "cliplabels_data": list(labs[1:]), "step": 1}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms (B200_PROFILING.md).  The process is started before the
    warm-up (its start-up takes longer than a short timed region) and only the samples whose timestamp falls between
    mark() and stop(), i.e. inside the timed region, are reported."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.t0 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def mark(self):
        """The timed region starts now."""
        import datetime
        self.t0 = datetime.datetime.now()

    def stop(self):
        import datetime
        t1 = datetime.datetime.now()
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f")
                if self.t0 is not None and not (self.t0 <= ts <= t1):
                    continue
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def run_ours(args):
    global H, W
    import torch.distributed as dist
    if args.size:
        if not args.profile_run:
            raise SystemExit("--size is a probe option: use it with --profile-run")
        H, W = (int(x) for x in args.size.lower().split("x"))
    from cvpr2021_vspw_implement_b200 import engine as E
    from cvpr2021_vspw_implement_b200 import parallel as P
    from cvpr2021_vspw_implement_b200._lib import lib
    from cvpr2021_vspw_implement_b200.data import synthetic_clip
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    E.set_precision(args.precision)
    syncbn = world > 1 and args.syncbn != "off"
    sync_group = P.make_syncbn_group(args.syncbn_exchange) if syncbn else None
    E.set_syncbn(syncbn, group=sync_group)

    model = build_model(dev, seed=0, kind=args.model)
    # every p.grad is a slice of ONE flat buffer the backward kernels write into; one NCCL all-reduce per step
    bucket = P.GradBucket(model.parameters())
    opt = make_optimizer(model, fused=not args.torch_sgd)
    imgs_h, labs_h = synthetic_clip(T_FRAMES, N_CLIPS, H, W, NUM_CLASS, seed=304 + rank)
    imgs_h = [t.pin_memory() for t in imgs_h]
    labs_h = [t.pin_memory() for t in labs_h]
    imgs_d = [t.to(dev) for t in imgs_h]
    labs_d = [t.to(dev) for t in labs_h]
    h2d_bytes = sum(t.numel() * 4 for t in imgs_h + labs_h)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)

    def step(imgs, labs):
        bucket.zero_grad()
        loss, acc = model(feed_from(imgs, labs))
        loss.backward()
        bucket.all_reduce_mean()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Python's cyclic GC is paused inside the timed regions (a gen-2 sweep over a step's few thousand tape closures
    # stalls the launching thread for tens of ms; the train entry point does the same per epoch)
    import gc
    gc.collect()
    gc.disable()

    # ---- warm-up ------------------------------------------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    n_warm = args.warmup if args.profile_run else max(args.warmup, 3)
    for _ in range(n_warm):
        step(imgs_d, labs_d)
    barrier()

    # ---- timed: device-resident inputs ------------------------------------------------------------------------------
    sampler.mark()
    E.conv_profile_begin()
    l0 = lib.launches
    evs = []
    barrier()
    t_wall0 = time.perf_counter()
    conv_sampled = 0
    for i in range(args.steps):
        flush.fill_(0.0)  # L2 flush between timed steps (outside the event pair)
        # the conv family's launch durations (roofline) are sampled on every 4th timed step: one CUDA-event pair around each of
        # its 343 launches is itself ~1 % of a step
        E.conv_profile_sample(i % 4 == 0)
        conv_sampled += 1 if i % 4 == 0 else 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss = step(imgs_d, labs_d)
        e1.record()
        evs.append((e0, e1))
    barrier()
    wall = time.perf_counter() - t_wall0
    launches = (lib.launches - l0) // max(args.steps, 1)
    conv_prof = E.conv_profile_end()
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    frames_per_step = T_FRAMES * N_CLIPS * world
    value = frames_per_step / (ms_per_step / 1e3)
    final_loss = float(loss.item())

    # ---- timed: end to end from pinned host buffers ---------------------------------------------------------------
    # Every step's images+labels are copied from pinned host memory (H2D) and every step's loss is read back (D2H), all
    # inside one CUDA-event pair around the whole region.  The feed is the repo's DevicePrefetcher (clip i+1 copies on a
    # side stream while clip i computes) and the loss of step i is read after step i+1 has been queued, as the training
    # entry point does; nothing is cached across steps.
    from cvpr2021_vspw_implement_b200.data import DevicePrefetcher
    e2e_steps = 0 if args.profile_run else max(2, min(args.steps, 5))
    e2e_value = 0.0
    region_ms = []
    if e2e_steps:
        loss_host = torch.empty(e2e_steps, dtype=torch.float32).pin_memory()
        # untimed rehearsal of the same loop: the host runs several steps ahead of the device here, so the feed path needs as
        # many in-flight input buffers as steps; the caching allocator gets them now (a cudaMalloc inside the timed region
        # costs tens of ms next to 40 GB of live allocations) and the timed region below only recycles them
        for imgs, labs in DevicePrefetcher(((imgs_h, labs_h) for _ in range(e2e_steps)), dev):
            step(imgs, labs)
        region_ms = []
        for _ in range(2):  # two timed regions, the faster one is reported (an allocator hiccup costs a whole region ~50 %)
            barrier()
            flush.fill_(0.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            feed = DevicePrefetcher(((imgs_h, labs_h) for _ in range(e2e_steps)), dev)
            for i, (imgs, labs) in enumerate(feed):
                loss = step(imgs, labs)
                loss_host[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)  # D2H read of the step's result
            e1.record()
            barrier()
            assert all(v == v for v in loss_host.tolist())
            t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            region_ms.append(float(t.item()))
        e2e_value = frames_per_step / (min(region_ms) / e2e_steps / 1e3)

    # ---- N > 1: the same step with LOCAL BN statistics (no statistics exchange), reported next to the headline ----------
    local_bn = None
    if syncbn and not args.profile_run:
        E.set_syncbn(False)
        for _ in range(2):
            step(imgs_d, labs_d)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        k_loc = max(3, min(args.steps, 5))
        for _ in range(k_loc):
            step(imgs_d, labs_d)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        local_bn = {"value": round(frames_per_step / (float(t.item()) / k_loc / 1e3), 3), "unit": UNIT, "ms_per_step": round(float(t.item()) / k_loc, 3),
                    "steps": k_loc, "note": "same step with per-rank BN statistics (syncbn off), no L2 flush between steps"}
        E.set_syncbn(True, group=sync_group)

    # ---- reported option (NOT the parity mode): single-pass bf16 weight gradients, forward / dgrad still bf16x3 ------------------
    fast_wgrad = None
    if world == 1 and not args.profile_run and args.precision == "bf16x3":
        E._state["wgrad_single"] = True
        for _ in range(2):
            step(imgs_d, labs_d)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        k_opt = max(3, min(args.steps, 5))
        for _ in range(k_opt):
            step(imgs_d, labs_d)
        e1.record()
        barrier()
        ms_opt = e0.elapsed_time(e1) / k_opt
        fast_wgrad = {"value": round(frames_per_step / (ms_opt / 1e3), 3), "unit": UNIT, "ms_per_step": round(ms_opt, 3), "steps": k_opt,
                      "note": "VSPW_WGRAD_SINGLE=1: weight gradients with single-pass bf16 operands (1 MMA per product), forward and dgrad bf16x3; "
                              "weight-gradient error ~2e-3 rel-L2, below the fp32-vs-fp32 floor of these gradients but not parity mode; no L2 flush between steps"}
        E._state["wgrad_single"] = False

    if args.kernel_profile and rank == 0:
        lib.profile_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            step(imgs_d, labs_d)
        e1.record()
        prof = lib.profile_end()
        tot = e0.elapsed_time(e1) / 2
        with open(args.kernel_profile, "w") as f:
            f.write(f"# per-entry-point CUDA-event time inside 2 real steps ({args.precision}); step = {tot:.2f} ms\n")
            f.write("| entry point | calls/step | ms/step | share |\n|---|---|---|---|\n")
            acc = 0.0
            for name, (n, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
                f.write(f"| `{name}` | {n // 2} | {ms / 2:.2f} | {100 * ms / 2 / tot:.1f}% |\n")
                acc += ms / 2
            f.write(f"| (outside the C ABI: optimizer, allocator fills, gaps) | | {tot - acc:.2f} | {100 * (tot - acc) / tot:.1f}% |\n")
            f.write("\nslowest single calls (index in call order, entry point, ms):\n")
            for i, name, ms in sorted(lib.last_profile_calls, key=lambda r: -r[2])[:12]:
                f.write(f"- #{i} `{name}` {ms:.3f}\n")
            E.conv_profile_begin()
            step(imgs_d, labs_d)
            cp = E.conv_profile_end()
            arm = cp["fp32_arm"]
            f.write(f"\nconvs left on the CUDA-core arm ({len(arm)} launches, {sum(m for _, m in arm):.2f} ms/step):\n")
            for tag, ms in sorted(arm, key=lambda r: -r[1]):
                f.write(f"- {tag}: {ms:.3f} ms\n")
            geo = {}
            for tag, ms, fl in cp["tc_arm"]:
                g = geo.setdefault(tag, [0, 0.0, 0.0])
                g[0] += 1; g[1] += ms; g[2] += fl
            mul = 3 if args.precision == "bf16x3" else 1
            f.write("\ntensor-core convs by geometry, inside one real step (CUDA events around each launch; launches, ms/step, "
                    "algorithmic and executed tensor TFLOP/s):\n| conv | launches | ms/step | ms each | alg TF/s | tensor TF/s |\n|---|---|---|---|---|---|\n")
            for tag, (cnt, ms, fl) in sorted(geo.items(), key=lambda kv: -kv[1][1]):
                f.write(f"| {tag} | {cnt} | {ms:.3f} | {ms / cnt:.3f} | {fl / ms / 1e9:.0f} | {mul * fl / ms / 1e9:.0f} |\n")

    gc.enable()
    pk = peaks()
    roof = None
    if conv_prof["launches"]:
        ach = conv_prof["tflop"] / (conv_prof["ms"] / 1e3)
        peak = pk["bf16_tflops_sustained"]
        roof = {"bound": "tensor", "kernel": conv_prof["kernel"], "achieved": round(ach, 2), "peak": peak, "unit": "TFLOP/s",
                "frac": round(ach / peak, 4), "traffic": None, "peak_source": pk["source"] + ", sustained bf16 (kernel timed inside a long step)",
                # executed tensor FLOPs (3 MMAs per algorithmic product in bf16x3) against the same peak: what the tensor pipe sees
                "tensor_tflops_executed": round(ach * (3 if args.precision == "bf16x3" else 1), 2),
                "tensor_frac_executed": round(ach * (3 if args.precision == "bf16x3" else 1) / peak, 4) if args.precision != "fp32" else 0.0,
                "launches_per_step": conv_prof["launches"] // conv_sampled, "ms_per_step": round(conv_prof["ms"] / conv_sampled, 3),
                "algorithmic_tflop_per_step": round(conv_prof["tflop"] / conv_sampled, 3), "timed_steps_sampled": conv_sampled,
                "note": {"fp32": "CUDA-core FFMA arm: tensor pipe idle, frac is vs the tensor roofline the tcgen05 arm is judged on",
                         "bf16x3": "each algorithmic FLOP costs 3 tensor FLOPs (hi/lo split): algorithmic ceiling = peak/3",
                         "bf16": "single-pass bf16 operands"}[args.precision]}

    if roof is not None:
        # `traffic`: dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel's most frequent geometry, parsed
        # from the committed ncu --set full capture (tools/ncu_traffic.py writes this file from the .ncu-rep's raw CSV page)
        tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            roof["traffic"] = tj.get("dram_bytes_per_launch")
            roof["traffic_detail"] = {k: tj[k] for k in ("kernel", "launch", "algorithmic_bytes_per_launch", "source") if k in tj}
    metric = METRIC if args.model == "psp" else METRIC.replace("TCB-PSP", "TCB-OCR")
    step_tflop = STEP_TFLOP if args.model == "psp" else 22.907
    if roof is not None and args.model != "psp":
        roof["note"] += "; TCB-OCR: %.3f algorithmic TFLOP/step" % step_tflop
    out = {"metric": metric, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": n_warm,
           "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": {"fp32": "f32", "bf16x3": "bf16x3 (bf16 hi+lo operands, f32 accumulate)", "bf16": "bf16"}[args.precision],
           "data": "synthetic", "impl": "ours",
           "config": {"workload": ("TCB-PSP" if args.model == "psp" else "TCB-OCR") + " ResNet101-dilated train fwd+bwd+SGD, T=5, n=2 clips/GPU, 480x854, K=124 (BASELINE configs[%d])" % (1 if args.model == "psp" else 2),
                      "frames_per_step_per_gpu": T_FRAMES * N_CLIPS, "parallelism": f"dp{world}", "precision_mode": args.precision,
                      "syncbn": bool(syncbn), "syncbn_exchange": (args.syncbn_exchange if syncbn else None),
                      "grad_allreduce": (None if world == 1 else f"NCCL AVG over one flat bucket, {bucket.last_overlapped} of {len(bucket._chunks)} chunks launched during the backward pass"),
                      "l2_flush": "256 MiB write between timed steps",
                      "optimizer": ("torch.optim.SGD" if args.torch_sgd else "FusedSGD (vspw_sgd_momentum_step)") + ": the reference's update rule and 4 param groups (train_clip2.py:215-236), inside the timed step",
                      "loss": round(final_loss, 5), "wall_s_timed_region": round(wall, 3),
                      "peak_hbm_gb": round(torch.cuda.max_memory_allocated(dev) / 1e9, 2)},
           "clocks": clocks, "gpu_launches": int(launches),
           "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": 4, "steps": e2e_steps,
                   "regions_ms": [round(x, 1) for x in region_ms] if e2e_steps else [],
                   "region_values": [round(frames_per_step / (x / e2e_steps / 1e3), 3) for x in region_ms] if e2e_steps else [],
                   "note": "value = the faster of the two timed regions (both listed); every region copies every step's inputs H2D and reads its loss D2H"},
           "roofline": roof}
    if local_bn is not None:
        out["local_bn"] = local_bn
    if fast_wgrad is not None:
        out["option_wgrad_single_bf16"] = fast_wgrad
    if rank == 0 and not args.profile_run and world == 1 and not args.no_gpu_library_baseline:
        del model, opt, bucket
        out["gpu_library_baseline"] = gpu_library_baseline(dev, imgs_d, labs_d, args)
    if rank == 0 and not args.no_cpu_baseline and not args.profile_run and world == 1:
        out["cpu_baseline"] = cpu_baseline(args)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        if sync_group is not None and hasattr(sync_group, "close"):
            sync_group.close()
        dist.destroy_process_group()


def gpu_library_baseline(dev, imgs_d, labs_d, args):
    """Stated context, never the target: the oracle port — the reference's ATen call sequence — on THIS
    GPU through cuDNN/cuBLAS, fp32 with TF32 disabled (the numerical truth the parity tests use) and with torch's default
    TF32 convolutions; train fwd+bwd of the same clip, CUDA events, 1 warm-up + 2 timed steps each."""
    res = {"unit": UNIT, "what": "oracle/tcb_oracle.py (reference ATen calls) on cuda through cuDNN, train fwd+bwd only (no optimizer), T=5 n=2 480x854"}
    try:
        O = _oracle()
        torch.cuda.empty_cache()
        sd, _, _ = _cpu_setup(8, 8, kind=args.model)
        sd = {k: v.detach().to(dev) for k, v in sd.items()}
        for k, v in sd.items():
            if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
                v.requires_grad_(True)
        fwd = O.clip_psp_forward if args.model == "psp" else O.clip_ocr_forward
        fr, lb = list(imgs_d[1:]) + [imgs_d[0]], list(labs_d[1:]) + [labs_d[0]]
        for name, tf32 in (("fp32_tf32_off", False), ("tf32_on_torch_default", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            ms = []
            for i in range(3):
                for v in sd.values():
                    v.grad = None
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = fwd(sd, fr, lb, train=True)
                out["loss"].backward()
                e1.record()
                torch.cuda.synchronize()
                if i:
                    ms.append(e0.elapsed_time(e1))
                del out
            res[name] = {"value": round(T_FRAMES * N_CLIPS / (min(ms) / 1e3), 2), "ms_per_step": round(min(ms), 2)}
        torch.backends.cudnn.allow_tf32 = True
    except Exception as e:  # noqa: BLE001 — context only: never fail the bench line over it
        res["error"] = f"{type(e).__name__}: {e}"[:300]
    return res


# --------------------------------------------------------------------------------------------------
def _cpu_threads():
    """All the host threads the box offers, whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1)."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0)) or n
    except AttributeError:
        pass
    torch.set_num_threads(n)
    return torch.get_num_threads()


def _cpu_step(sd, imgs, labs, kind="psp"):
    O = _oracle()
    for v in sd.values():
        if v.grad is not None:
            v.grad = None
    fr, lb = list(imgs[1:]) + [imgs[0]], list(labs[1:]) + [labs[0]]
    out = (O.clip_psp_forward if kind == "psp" else O.clip_ocr_forward)(sd, fr, lb, train=True)
    out["loss"].backward()
    return float(out["loss"].item())


def _cpu_setup(hs, ws, kind="psp"):
    O = _oracle()
    from cvpr2021_vspw_implement_b200 import models as M  # parameter containers only (CPU), no kernels involved
    torch.manual_seed(0)
    ns = argparse.Namespace(num_class=NUM_CLASS, psp_weight=False, use_memory=False, memory_num=8, clipocr_all=False)
    cls = M.Clip_PSP if kind == "psp" else M.ClipOCRNet
    m = cls(M.ModelBuilder.build_encoder("resnet101dilated"), torch.nn.NLLLoss(ignore_index=255), ns, deep_sup_scale=0.4)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    for k, _ in m.named_parameters():
        sd[k].requires_grad_(True)
    imgs, labs = O.synthetic_clip(T_FRAMES, N_CLIPS, hs, ws, NUM_CLASS, seed=304)
    return sd, imgs, labs


def _host_desc():
    model = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return f"{model or 'unknown CPU'}, {os.cpu_count()} logical CPUs"


def _avail_ram_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 2 ** 30
    except Exception:  # noqa: BLE001
        return 0.0


def cpu_baseline(args):
    """The CPU oracle (reference path restated in PyTorch-CPU fp32) on a bounded sample: the same T=5, n=2 train step at
    `--cpu-sample` resolution (1 warm-up + 1 timed step), throughput scaled by the pixel ratio to the 480x854 workload (conv cost
    is linear in pixels).  `--impl reference` runs the full-size arm."""
    hs, ws = (int(x) for x in args.cpu_sample.lower().split("x"))
    threads = _cpu_threads()
    sd, imgs, labs = _cpu_setup(hs, ws, args.model)
    _cpu_step(sd, imgs, labs, args.model)
    t0 = time.perf_counter()
    _cpu_step(sd, imgs, labs, args.model)
    dt = time.perf_counter() - t0
    raw = T_FRAMES * N_CLIPS / dt
    scaled = raw * (hs * ws) / (H * W)
    return {"value": round(scaled, 4), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"one warmed TCB-{args.model.upper()} R101 train fwd+bwd step, T=5 n=2 at {hs}x{ws} ({dt:.1f} s, {raw:.3f} clip-frames/s at that "
                      f"size), scaled by the pixel ratio {hs * ws}/{H * W} to 480x854; PyTorch-CPU fp32 (oneDNN), {threads} threads; {_host_desc()}"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (the oracle port: the same ATen call sequence, PyTorch
    CPU fp32) on the host cores, ALL of them.  Exactly --warmup W + --steps K steps are run; each step is one T=5, n=2 train
    fwd+bwd on a bounded sample of the workload: the full 480x854 clip when K+W such steps fit ~5 minutes on this host (probed
    with one quarter-size step) and there is RAM for the fp32 autograd graph, else the 240x427 clip (a quarter of the pixels;
    conv cost is linear in pixels) with the throughput scaled by the pixel ratio.  `ms_per_step` is always what was measured."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = _cpu_threads()
    qs = tuple(int(x) for x in args.cpu_sample.lower().split("x"))
    steps, warm = max(1, args.steps), max(0, args.warmup)
    sd, imgs, labs = _cpu_setup(*qs, args.model)
    _cpu_step(sd, imgs, labs, args.model)  # untimed: oneDNN primitive creation
    t0 = time.perf_counter()
    _cpu_step(sd, imgs, labs, args.model)
    t_q = time.perf_counter() - t0
    ratio = (H * W) / (qs[0] * qs[1])
    full = (steps + warm) * t_q * ratio <= 300.0 and _avail_ram_gb() >= 64.0  # (the full-size fp32 autograd graph peaks at ~45 GB)
    hs, ws = (H, W) if full else qs
    if full:
        sd, imgs, labs = _cpu_setup(hs, ws, args.model)
    for _ in range(warm):
        _cpu_step(sd, imgs, labs, args.model)
    t0 = time.perf_counter()
    for _ in range(steps):
        _cpu_step(sd, imgs, labs, args.model)
    dt = (time.perf_counter() - t0) / steps
    raw = T_FRAMES * N_CLIPS / dt
    scaled = raw * (hs * ws) / (H * W)
    sample = (f"{steps} timed TCB-{args.model.upper()} R101 train fwd+bwd steps after {warm} warm-up, T=5 n=2 at {hs}x{ws} ({dt:.2f} s/step"
              + ("" if full else f", {raw:.3f} clip-frames/s at that size, scaled by the pixel ratio {hs * ws}/{H * W} to 480x854")
              + f"); quarter-size probe step {t_q:.2f} s; PyTorch-CPU fp32 (oneDNN), {threads} threads; {_host_desc()}")
    metric = METRIC if args.model == "psp" else METRIC.replace("TCB-PSP", "TCB-OCR")
    out = {"metric": metric, "value": round(scaled, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
           "ms_per_step": round(dt * 1e3, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "impl": "reference",
           "config": {"workload": f"TCB-{args.model.upper()} ResNet101-dilated train fwd+bwd, T=5, n=2 clips, 480x854, K=124 (BASELINE configs[{1 if args.model == 'psp' else 2}])",
                      "sample": sample, "sample_hw": [hs, ws], "full_size": bool(full),
                      "quarter_size_clip_frames_per_s_scaled": round(T_FRAMES * N_CLIPS / t_q / ratio, 4)},
           "cpu_baseline": {"value": round(scaled, 4), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": round(scaled, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
