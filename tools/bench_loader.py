#!/usr/bin/env python
"""Loader throughput: can the data path into the boundary keep up with the engine?

    python tools/bench_loader.py [--workers W] [--videos V] [--frames F] [--size 480x854] [--cropsize 479] [--multi-scale]

Generates a VSPW-shaped directory (JPEG frames + PNG masks at `--size`), then drives `vspw_data.VSPWClipTrain` (= the
reference's `BaseDataset_longclip`, dataset2.py:852-1048: JPEG/PNG decode, clip sampling, mirror, multi-scale, pad + shared
crop, label remap, ImageNet normalise) through a `DataLoader` with W worker processes and pinned batches, plus — when a GPU is
present — `data.DevicePrefetcher` (H2D on a side stream).  Prints one JSON line: clip-frames/s delivered, per worker count,
next to what one engine rank consumes (bench.py's `value`)."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cvpr2021_vspw_implement_b200.data import DevicePrefetcher  # noqa: E402
from cvpr2021_vspw_implement_b200.vspw_data import VSPWClipTrain  # noqa: E402


def make_tree(root, videos, frames, h, w, seed=0):
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    for v in range(videos):
        d = os.path.join(root, "data", f"vid_{v:03d}")
        os.makedirs(os.path.join(d, "origin"))
        os.makedirs(os.path.join(d, "mask"))
        # photo-like content (smooth gradients + texture): uniform noise would overstate the JPEG decode cost
        base = np.stack([128 + 90 * np.sin(xx / rng.uniform(20, 90) + c) * np.cos(yy / rng.uniform(20, 90)) for c in range(3)], -1)
        lab = (rng.randint(0, 125, ((h + 31) // 32, (w + 31) // 32)).repeat(32, 0).repeat(32, 1)[:h, :w]).astype(np.uint8)
        for f in range(frames):
            img = np.clip(base + rng.normal(0, 12, base.shape) + 3 * f, 0, 255).astype(np.uint8)
            Image.fromarray(img).save(os.path.join(d, "origin", f"{f:08d}.jpg"), quality=92)
            Image.fromarray(lab).save(os.path.join(d, "mask", f"{f:08d}.png"))
    for split in ("train", "val"):
        with open(os.path.join(root, split + ".txt"), "w") as fh:
            fh.write("".join(f"vid_{v:03d}\n" for v in range(videos)))


def measure(ds, workers, batch, iters, device, finish_u8=False):
    dl = torch.utils.data.DataLoader(ds, batch_size=batch, shuffle=True, num_workers=workers, drop_last=True, pin_memory=True,
                                     persistent_workers=workers > 0, prefetch_factor=4 if workers > 0 else None)
    done, t0, frames = 0, None, 0
    while done < iters + 2:
        it = DevicePrefetcher(dl, device, finish_u8=finish_u8) if device is not None else iter(dl)
        for imgs, labs in it:
            if done == 2:  # two untimed batches: worker start-up, first decode
                if device is not None:
                    torch.cuda.synchronize()
                t0 = time.perf_counter()
                frames = 0
            elif done > 2:
                frames += len(imgs) * imgs[0].shape[0]
            done += 1
            if done >= iters + 3:
                break
        else:
            continue
        break
    if device is not None:
        torch.cuda.synchronize()
    return frames / (time.perf_counter() - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", default="0,4,8,16")
    ap.add_argument("--videos", type=int, default=24)
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--size", default="480x854")
    ap.add_argument("--cropsize", type=int, default=479)
    ap.add_argument("--clip-num", type=int, default=5)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--iters", type=int, default=24)
    ap.add_argument("--multi-scale", action="store_true")
    ap.add_argument("--device-finish", action="store_true", help="uint8 items, float conversion / normalisation / label remap on the GPU")
    a = ap.parse_args()
    h, w = (int(x) for x in a.size.lower().split("x"))
    device = torch.device("cuda", 0) if torch.cuda.is_available() else None
    out = {"what": "VSPWClipTrain (= dataset2.BaseDataset_longclip) -> DataLoader(pin_memory) -> DevicePrefetcher", "frame_size": [h, w],
           "cropsize": a.cropsize, "clip_num": a.clip_num, "multi_scale": bool(a.multi_scale), "host_cpus": os.cpu_count(),
           "h2d": device is not None, "device_finish": bool(a.device_finish and device is not None), "clip_frames_per_s": {}}
    with tempfile.TemporaryDirectory() as root:
        make_tree(root, a.videos, a.frames, h, w)
        args = argparse.Namespace(cropsize=a.cropsize, dataroot=root, trainfps=1, clip_num=a.clip_num,
                                  dilation2=",".join(str(i + 1) for i in range(a.clip_num - 1)), multi_scale=a.multi_scale, lesslabel=False)
        fin = bool(a.device_finish and device is not None)
        ds = VSPWClipTrain(args, "train", device_finish=fin)
        for wk in [int(x) for x in a.workers.split(",")]:
            out["clip_frames_per_s"][str(wk)] = round(measure(ds, wk, a.batch, a.iters, device, fin), 1)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
