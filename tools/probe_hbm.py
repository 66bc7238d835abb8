#!/usr/bin/env python
"""HBM direction probe: write-only (fill), read-only (sum), copy on 1 GiB fp32 buffers, CUDA events, best of 10.
The wide-N / small-K 1x1 convs write 4 fp32 bytes per output element and read a quarter of that: their floor is the WRITE rate."""
import torch

dev = torch.device("cuda", 0)
n = 1 << 28  # 1 GiB of fp32
a = torch.empty(n, device=dev)
b = torch.empty(n, device=dev)
a.fill_(1.0); b.fill_(2.0)


def best(fn, k=10):
    ts = []
    for _ in range(k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


gb = n * 4 / 1e9
print(f"write-only  fill_ 1 GiB : {best(lambda: a.fill_(3.0)):.3f} ms  {gb / best(lambda: a.fill_(3.0)) * 1e3:.0f} GB/s written")
t = best(lambda: a.sum())
print(f"read-only   sum   1 GiB : {t:.3f} ms  {gb / t * 1e3:.0f} GB/s read")
t = best(lambda: b.copy_(a))
print(f"copy        1 GiB -> 1 GiB : {t:.3f} ms  {2 * gb / t * 1e3:.0f} GB/s read + written")
