#!/usr/bin/env python
"""Inference paths of the hot path, timed on one B200 with CUDA events and checked against the oracle:

  * BASELINE configs[0]: PSPNet ResNet18-dilated + PPMDeepsup, ONE 480x854 frame, batch 1, eval forward -> (1, 124, 480, 854)
    probabilities — our CUDA path, and the oracle (the reference's ATen calls) on this box's host cores beside it;
  * TCB-PSP / TCB-OCR ResNet101-dilated, T=5, n=2, eval forward with segSize=(480, 854).

    python tools/bench_eval.py [--iters 10] [--no-cpu]      -> markdown table on stdout
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import cases as C  # noqa: E402
import tcb_oracle as O  # noqa: E402
from cvpr2021_vspw_implement_b200 import engine as E  # noqa: E402

H, W, K = 480, 854, 124


def gpu_ms(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    rows = []
    # ---- configs[0]: R18-dilated + PPMDeepsup, one frame -----------------------------------------------------------------
    m = C.build("SegmentationModule", "resnet18dilated", 14).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    imgs, labs = O.synthetic_clip(1, 1, H, W, K, seed=307, block=16)
    img = imgs[0]
    cpu_s = None
    with torch.no_grad():
        if not a.no_cpu:
            torch.set_num_threads(os.cpu_count() or 1)
            O.segmentation_module_forward(sd, img, None, train=False, seg_size=(H, W))
            t0 = time.perf_counter()
            ref = O.segmentation_module_forward(sd, img, None, train=False, seg_size=(H, W))
            cpu_s = time.perf_counter() - t0
            ref = ref["probs"] if isinstance(ref, dict) else ref
        mg = m.to(dev)
        xg = img.to(dev)
        for prec in ("bf16x3", "bf16"):
            with E.precision(prec):
                ms, out = gpu_ms(lambda: mg({"img_data": xg}, segSize=(H, W)), a.iters)
            err = agree = None
            if cpu_s is not None:
                err = C.rel_err(out.cpu(), ref)
                agree = float((out.argmax(1).cpu() == ref.argmax(1)).float().mean())
            rows.append((f"configs[0] PSPNet R18-dilated + PPMDeepsup, 1 x 480x854 frame, eval ({prec})", ms, 1000.0 / ms, err, agree))
        if cpu_s is not None:
            rows.append((f"  the oracle (reference ATen calls) on this box's {torch.get_num_threads()} host threads", cpu_s * 1e3, 1.0 / cpu_s, None, None))
    del mg
    # ---- TCB models, inference tail ----------------------------------------------------------------------------------------
    for kind, name in (("Clip_PSP", "TCB-PSP"), ("ClipOCRNet", "TCB-OCR")):
        m = C.build(kind, "resnet101dilated", 21).eval().to(dev)
        imgs, labs = O.synthetic_clip(5, 2, H, W, K, seed=304)
        feed = C.feed(imgs, labs, False, dev)
        with torch.no_grad():
            for prec in ("bf16x3", "bf16"):
                with E.precision(prec):
                    ms, out = gpu_ms(lambda: m(dict(feed, clipimgs_data=list(feed["clipimgs_data"])), segSize=(H, W)), a.iters)
                rows.append((f"{name} R101-dilated, T=5, n=2, 480x854, eval forward + up-sampled softmax ({prec})", ms, 10 * 1000.0 / ms, None, None))
        del m
        torch.cuda.empty_cache()
    print("| path | ms | frames/s (clip-frames/s for the TCB models) | probabilities max-abs vs oracle | argmax agreement |\n|---|---|---|---|---|")
    for what, ms, fps, err, agree in rows:
        print(f"| {what} | {ms:.2f} | {fps:.1f} | {'' if err is None else f'{err:.1e}'} | {'' if agree is None else f'{100 * agree:.3f} %'} |")


if __name__ == "__main__":
    main()
