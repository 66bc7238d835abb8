#!/usr/bin/env python
"""profiles/r2_traffic.json from an `ncu --set full` capture: DRAM bytes per launch of the dominant kernel's most frequent launch.

    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep --kernel conv_tc2_kernel --pick median \\
        --launch "layer3 3x3 d2 256->256 fwd (bf16x3)" --algorithmic-bytes 134.0e6 --out profiles/r2_traffic.json

bench.py copies `dram_bytes_per_launch` into `roofline.traffic` (B200_PROFILING.md: traffic = dram__bytes_read.sum +
dram__bytes_write.sum of one capture, per launch)."""
import argparse
import csv
import json
import statistics
import subprocess

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--kernel", required=True)
    ap.add_argument("--pick", default="median", help="median | first | index N (among the matching launches, by duration)")
    ap.add_argument("--launch", default="")
    ap.add_argument("--algorithmic-bytes", type=float, default=None)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    txt = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader([l for l in txt.splitlines() if not l.startswith("==")]))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name, table):
        v = float(r[col[name]].replace(",", ""))
        return v * table.get(units[col[name]], 1.0)

    sel = [r for r in rows[2:] if a.kernel in r[col["Kernel Name"]]]
    if not sel:
        raise SystemExit(f"no launch of {a.kernel!r} in {a.rep}")
    recs = []
    for r in sel:
        recs.append({"kernel": r[col["Kernel Name"]].split("(")[0].replace("<unnamed>::", "").replace("void ", ""),
                     "duration_us": val(r, "gpu__time_duration.sum", TIME),
                     "dram_read": val(r, "dram__bytes_read.sum", UNIT), "dram_write": val(r, "dram__bytes_write.sum", UNIT),
                     "tensor_pipe_pct": float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]) if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in col else None,
                     "dram_pct": float(r[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
                     "grid": r[col["Grid Size"]], "registers": r[col["launch__registers_per_thread"]]})
    recs.sort(key=lambda x: x["duration_us"])
    if a.pick == "median":
        pick = recs[len(recs) // 2]
    elif a.pick == "first":
        pick = recs[0]
    else:
        pick = recs[int(a.pick)]
    out = {"kernel": pick["kernel"], "launch": a.launch, "dram_bytes_per_launch": pick["dram_read"] + pick["dram_write"],
           "dram_read_bytes": pick["dram_read"], "dram_write_bytes": pick["dram_write"], "duration_us_under_ncu": pick["duration_us"],
           "tensor_pipe_pct": pick["tensor_pipe_pct"], "dram_pct_of_peak": pick["dram_pct"], "grid": pick["grid"], "registers": pick["registers"],
           "algorithmic_bytes_per_launch": a.algorithmic_bytes, "launches_in_capture": len(recs),
           "source": f"ncu --set full --clock-control none, {a.rep.split('/')[-1]} (cold L2: ncu flushes caches between replays)"}
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
