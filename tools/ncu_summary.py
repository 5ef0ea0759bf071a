#!/usr/bin/env python
"""Summarise ncu outputs into the small text files kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_r1.csv > profiles/r1_launches.txt
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep > profiles/r1_full_<kernel>.txt
"""
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors.sum", "lts__t_sectors_srcunit_tex.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_warps", "sm__maximum_warps_per_active_cycle_pct",
    "sm__cycles_elapsed.avg.per_second", "lts__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg, order = {}, []
    for r in rows:
        if not r[0].isdigit():
            continue
        name = re.sub(r"\(.*", "", r[ki])
        if name not in agg:
            order.append(name)
        agg.setdefault(name, []).append(float(r[vi].replace(",", "")))
    total = sum(sum(v) for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none  (cold-cache, serialised: compare SHARES)")
    print(f"# {'kernel':70s} {'n':>4s} {'avg ms':>10s} {'total ms':>10s} {'share':>7s}")
    for k in order:
        v = agg[k]
        print(f"{k[:72]:72s} {len(v):4d} {sum(v) / len(v) / 1e6:10.4f} {sum(v) / 1e6:10.3f} {100 * sum(v) / total:6.1f}%")


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none: {path}")
    for r in rows[2:]:
        print(f"\n## {r[hdr.index('Kernel Name')]}  (launch id {r[0]})")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:80s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
