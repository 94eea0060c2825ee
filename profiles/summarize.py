#!/usr/bin/env python
"""Summarise ncu outputs into small text files that can be committed under profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv           > profiles/<round>_launches.txt
  python profiles/summarize.py kernels  gpurun_out/prof.ncu-rep           > profiles/<round>_kernels.txt
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"])
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(row["Metric Unit"], v)
        agg.setdefault(row["Kernel Name"].split("(")[0], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n# source: {path}")
    print(f"{'kernel':45s} {'launches':>8s} {'avg_us':>10s} {'total_us':>11s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:45s} {len(v):8d} {sum(v)/len(v):10.2f} {sum(v):11.1f} {sum(v)/tot*100:6.1f}%")


def kernels(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none --import-source on; source: {path}")
    for row in rows[2:]:
        print(f"\n== {row[hdr.index('Kernel Name')]}  (launch id {row[hdr.index('ID')]})")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {m:80s} {row[i]:>16s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "kernels": kernels}[sys.argv[1]](sys.argv[2])
