#!/usr/bin/env python
"""Per-geometry timing of the tcgen05 convolution kernels (fwd with/without fused BN statistics, dgrad, wgrad) on the
conv shapes of ResNet101-dilated TCB-PSP at 480x854, T=5, n=2.  CUDA events, L2 flushed between
launches.  Prints one line per (geometry, kernel): ms, algorithmic TFLOP/s, tensor TFLOP/s (x3 in bf16x3 mode).

    python tools/bench_conv.py [--precision bf16x3|bf16] [--iters 5] [--json out.json]
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cvpr2021_vspw_implement_b200 import engine as E  # noqa: E402
from cvpr2021_vspw_implement_b200._lib import ConvDesc, PREC_BF16, PREC_BF16X3, lib  # noqa: E402

# (name, n, h, w, cin, cout, k, dil, count in the network)
SHAPES = [
    ("l1 1x1 128->64", 10, 120, 214, 128, 64, 1, 1, 1),
    ("l1 3x3 64->64", 10, 120, 214, 64, 64, 3, 1, 3),
    ("l1 1x1 64->256", 10, 120, 214, 64, 256, 1, 1, 3),
    ("l1 1x1 256->64", 10, 120, 214, 256, 64, 1, 1, 2),
    ("l2 1x1 256->128", 10, 120, 214, 256, 128, 1, 1, 1),
    ("l2 3x3 128->128", 10, 60, 107, 128, 128, 3, 1, 3),
    ("l2 1x1 128->512", 10, 60, 107, 128, 512, 1, 1, 4),
    ("l2 1x1 512->128", 10, 60, 107, 512, 128, 1, 1, 3),
    ("l3 1x1 512->256", 10, 60, 107, 512, 256, 1, 1, 1),
    ("l3 3x3 256->256 d2", 10, 60, 107, 256, 256, 3, 2, 23),
    ("l3 1x1 256->1024", 10, 60, 107, 256, 1024, 1, 1, 23),
    ("l3 1x1 1024->256", 10, 60, 107, 1024, 256, 1, 1, 22),
    ("l3 1x1 512->1024 ds", 10, 60, 107, 512, 1024, 1, 1, 1),
    ("l4 1x1 1024->512", 10, 60, 107, 1024, 512, 1, 1, 1),
    ("l4 3x3 512->512 d4", 10, 60, 107, 512, 512, 3, 4, 3),
    ("l4 1x1 512->2048", 10, 60, 107, 512, 2048, 1, 1, 3),
    ("l4 1x1 2048->512", 10, 60, 107, 2048, 512, 1, 1, 2),
    ("l4 1x1 1024->2048 ds", 10, 60, 107, 1024, 2048, 1, 1, 1),
    ("deepsup 3x3 1024->512", 10, 60, 107, 1024, 512, 3, 1, 1),
    ("ppm 3x3 4096->512", 2, 60, 107, 4096, 512, 3, 1, 1),
    ("stem 3x3 64->64", 10, 240, 427, 64, 64, 3, 1, 1),
    ("stem 3x3 64->128", 10, 240, 427, 64, 128, 3, 1, 1),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "bf16"])
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--json", default="")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    prec = PREC_BF16X3 if a.precision == "bf16x3" else PREC_BF16
    x3 = a.precision == "bf16x3"
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    rows = []
    tot = {"fwd": 0.0, "fwd+stats": 0.0, "dgrad": 0.0, "wgrad": 0.0}
    for name, n, h, w, cin, cout, k, dil, count in SHAPES:
        if a.only and a.only not in name:
            continue
        pad = dil * (k - 1) // 2
        d = ConvDesc(n, h, w, cin, cout, k, k, 1, pad, dil, h, w, prec)
        assert lib.tc_supported(d), name
        g = torch.Generator(device=dev).manual_seed(0)
        mk = lambda *s: torch.randn(*s, device=dev, generator=g).to(torch.bfloat16)
        xh, xl = mk(n, h, w, cin), (mk(n, h, w, cin) if x3 else None)
        wh, wl = mk(cout, k, k, cin), (mk(cout, k, k, cin) if x3 else None)
        th, tl = mk(cin, k, k, cout), (mk(cin, k, k, cout) if x3 else None)
        gh, gl = mk(n, h, w, cout), (mk(n, h, w, cout) if x3 else None)
        y = torch.empty(n, h, w, cout, device=dev)
        dx = torch.empty(n, h, w, cin, device=dev)
        dw = torch.empty(cout, k, k, cin, device=dev)
        stats = torch.zeros(2, cout, device=dev, dtype=torch.float64)
        flops = 2.0 * n * h * w * cout * k * k * cin
        calls = {
            "fwd": lambda: lib.call("vspw_conv2d_fwd_tc", ctypes.byref(d), P(xh), P(xl), P(wh), P(wl), None, P(y), None, None, st),
            "fwd+stats": lambda: lib.call("vspw_conv2d_fwd_tc", ctypes.byref(d), P(xh), P(xl), P(wh), P(wl), None, P(y), P(stats[0]), P(stats[1]), st),
            "dgrad": lambda: lib.call("vspw_conv2d_dgrad_tc", ctypes.byref(d), P(gh), P(gl), P(th), P(tl), P(dx), 0, st),
            "wgrad": lambda: lib.call("vspw_conv2d_wgrad_tc", ctypes.byref(d), P(xh), P(xl), P(gh), P(gl), P(dw), st),
        }
        for kind, fn in calls.items():
            fn()
            torch.cuda.synchronize()
            ms = []
            for _ in range(a.iters):
                flush.fill_(0.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            m = sorted(ms)[len(ms) // 2]
            tf = flops / m / 1e9
            tot[kind] += m * count
            rows.append({"shape": name, "kernel": kind, "ms": round(m, 4), "alg_tflops": round(tf, 1), "tensor_tflops": round(tf * (3 if x3 else 1), 1), "count": count})
            print(f"{name:24s} {kind:10s} {m:8.3f} ms  {tf:7.1f} alg TF/s  {tf * (3 if x3 else 1):7.1f} tensor TF/s  x{count}", flush=True)
    print("network totals (ms, weighted by count):", {k: round(v, 2) for k, v in tot.items()})
    if a.json:
        json.dump({"precision": a.precision, "rows": rows, "totals_ms": tot}, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
