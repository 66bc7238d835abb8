#!/usr/bin/env python
"""Turn ncu outputs into the markdown summaries committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv            # per-kernel shares of a launch list
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep [more.ncu-rep] # key metrics of --set full captures
"""
import collections
import csv
import re
import subprocess
import sys

FULL_METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        ms = v / 1e6 if unit.startswith("n") else (v / 1e3 if unit.startswith("u") else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")[:80]
        agg[name][0] += 1
        agg[name][1] += ms
        tot += ms
    print(f"{sum(a[0] for a in agg.values())} launches, {tot:.1f} ms summed device time\n")
    print("| kernel | launches | ms | share |\n|---|---|---|---|")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"| `{k}` | {n} | {ms:.2f} | {100 * ms / tot:.2f}% |")


def full(paths):
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units, data = rows[0], rows[1], rows[2:]
        print(f"### {path.split('/')[-1]}\n")
        cols = []
        for key, label in FULL_METRICS:
            idx = [i for i, h in enumerate(hdr) if h == key]
            if idx:
                cols.append((label, idx[0]))
        ki = hdr.index("Kernel Name")
        print("| kernel | " + " | ".join(f"{l} ({units[i]})" if units[i] else l for l, i in cols) + " |")
        print("|---|" + "---|" * len(cols))
        for r in data:
            name = re.sub(r"\(.*", "", r[ki]).replace("<unnamed>::", "").replace("void ", "")[:40]
            print(f"| `{name}` | " + " | ".join(r[i] for _, i in cols) + " |")
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2:])
