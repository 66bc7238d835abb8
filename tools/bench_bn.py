#!/usr/bin/env python
"""Per-shape timing of the three BN passes (vspw_bn_train_fwd, vspw_bn_bwd_reduce, vspw_bn_bwd_apply) on the tensor shapes of
ResNet101-dilated TCB-PSP at 480x854, T=5, n=2, through the engine (`batchnorm_act` + its recorded backward), CUDA events per
entry point, L2 flushed between layers.  Prints ms, algorithmic bytes and GB/s per pass, and the count-weighted network total.

    python tools/bench_bn.py [--iters 5]            # honours VSPW_RELU_BITS=0 for A/B
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cvpr2021_vspw_implement_b200 import engine as E  # noqa: E402
from cvpr2021_vspw_implement_b200._lib import lib  # noqa: E402
from cvpr2021_vspw_implement_b200.models.sync_batchnorm import BatchNorm2d  # noqa: E402

# (name, n, h, w, c, residual, fp32_out, count)
SHAPES = [
    ("stem 64ch @240x427", 10, 240, 427, 64, False, True, 2),
    ("stem 128ch @240x427", 10, 240, 427, 128, False, True, 1),
    ("l1 interior 64ch @120x214", 10, 120, 214, 64, False, False, 6),
    ("l1 out 256ch @120x214 (+res)", 10, 120, 214, 256, True, True, 3),
    ("l2 interior 128ch @60x107", 10, 60, 107, 128, False, False, 8),
    ("l2 out 512ch @60x107 (+res)", 10, 60, 107, 512, True, True, 4),
    ("l3 interior 256ch @60x107", 10, 60, 107, 256, False, False, 46),
    ("l3 out 1024ch @60x107 (+res)", 10, 60, 107, 1024, True, True, 23),
    ("l4 interior 512ch @60x107", 10, 60, 107, 512, False, False, 6),
    ("l4 out 2048ch @60x107 (+res)", 10, 60, 107, 2048, True, True, 3),
    ("head 512ch @60x107 (10 frames)", 10, 60, 107, 512, False, True, 1),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    E.set_precision("bf16x3")
    tot = {"vspw_bn_train_fwd": 0.0, "vspw_bn_bwd_reduce": 0.0, "vspw_bn_bwd_apply": 0.0}
    print(f"relu bits: {os.environ.get('VSPW_RELU_BITS', '1') != '0'}")
    for name, n, h, w, c, res, fp32_out, count in SHAPES:
        g = torch.Generator(device=dev).manual_seed(c)
        y = torch.randn(n, h, w, c, generator=g, device=dev)
        r = torch.randn(n, h, w, c, generator=g, device=dev) if res else None
        go = torch.randn(n, h, w, c, generator=g, device=dev)
        bn = BatchNorm2d(c).to(dev).train()
        acc = {k: 0.0 for k in tot}
        for it in range(a.iters + 1):
            tape = E.Tape(True)
            yv = E.Var(y, needs_grad=True)
            yv.wants_grad_planes = True       # as when a tensor-core conv produced y
            yv.wants_grad_fp32 = False
            rv = E.Var(r, needs_grad=True) if res else None
            flush.fill_(0.0)
            lib.profile_begin()
            ov = E.batchnorm_act(tape, yv, bn, relu=True, residual=rv, training=True, fp32_out=fp32_out)
            ov.grad = go.clone()
            flush.fill_(0.0)
            tape.backward()
            prof = lib.profile_end()
            if it:
                for k in acc:
                    acc[k] += prof[k][1] / a.iters
        elems = n * h * w * c
        line = f"{name:34s} x{count:<3d}"
        for k, tag in (("vspw_bn_train_fwd", "fwd"), ("vspw_bn_bwd_reduce", "reduce"), ("vspw_bn_bwd_apply", "apply")):
            line += f"  {tag} {acc[k] * 1e3:7.1f} us"
            tot[k] += acc[k] * count
        line += f"   ({elems / 1e6:.1f} M elements)"
        print(line, flush=True)
    print("network totals (ms, weighted by count): " + ", ".join(f"{k.replace('vspw_bn_', '')} {v:.2f}" for k, v in tot.items())
          + f", sum {sum(tot.values()):.2f}")


if __name__ == "__main__":
    main()
