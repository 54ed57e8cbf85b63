#!/usr/bin/env python
"""Summarise an ncu report (read here with `ncu -i`, no GPU needed) into a small CSV for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_prof_summary.csv
    python tools/ncu_summary.py --launches gpurun_out/launches.csv profiles/r1_launches_summary.csv
"""
import collections
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]


def full(rep, out):
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"]).decode()
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index(m) for m in METRICS if m in hdr]
    kn = hdr.index("Kernel Name")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + ["%s [%s]" % (hdr[c], units[c]) for c in cols])
        for r in rows[2:]:
            w.writerow([r[kn].split("(")[0]] + [r[c] for c in cols])


def launches(src, out):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        agg.setdefault(row["Kernel Name"].split("(")[0], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "mean_ms", "total_ms", "share"])
        for k, v in agg.items():
            w.writerow([k, len(v), "%.5f" % (sum(v) / len(v)), "%.4f" % sum(v), "%.4f" % (sum(v) / tot)])


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[1], sys.argv[2])
