#!/usr/bin/env python
"""SASS evidence for the tensor-core kernels of libvspw_b200.so -> profiles/r2_sass_opcodes.txt: per kernel, the count of the
Blackwell-native opcodes (UTC*MMA = tcgen05.mma, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce-add, LDTM = tcgen05.ld,
UTCBAR = tcgen05.commit, SYNCS...TRYWAIT = mbarrier.try_wait)."""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "cvpr2021_vspw_implement_b200/csrc/libvspw_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pat = re.compile(r"\b(UTC[A-Z]*MMA[.A-Z0-9_]*|UTMALDG[.A-Z0-9_]*|UTMASTG[.A-Z0-9_]*|UTMAREDG[.A-Z0-9_]*|LDTM[.a-zA-Z0-9_]*|UTCBAR[.A-Z0-9_]*|SYNCS[.A-Z0-9_]*TRYWAIT[.A-Z0-9_]*)")
counts = collections.OrderedDict()
fn = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        dem = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        k = re.search(r"([A-Za-z_0-9]+_kernel)(<[^>]*>)?", dem)
        fn = (k.group(1) + (k.group(2) or "")) if k else dem[:70]
        continue
    m = pat.search(line)
    if m and fn:
        counts.setdefault(fn, collections.Counter())[m.group(1)] += 1
print("# SASS opcode histogram of the tensor-core kernels in libvspw_b200.so (cuobjdump -sass, sm_100a); regenerate: python tools/sass_opcodes.py")
print("# UTC*MMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce-add, LDTM = tcgen05.ld,")
print("# UTCBAR = tcgen05.commit, SYNCS...TRYWAIT = mbarrier.try_wait")
for fn, c in counts.items():
    if not any(op.startswith("UTC") for op in c):
        continue
    print(f"\n{fn}")
    for op, n in sorted(c.items()):
        print(f"    {n:4d}  {op}")
