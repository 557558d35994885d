#!/usr/bin/env python
"""Condense ncu outputs into the small text files kept under profiles/.

    tools/ncu_summary.py launches gpurun_out/launches_X.csv  > profiles/launches_X.txt
    tools/ncu_summary.py full gpurun_out/prof_X.ncu-rep      > profiles/ncu_X.txt
"""
import csv, io, subprocess, sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]

def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    agg = OrderedDict()
    total = 0.0
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu].strip(), 1.0)   # -> us
        name = r[ik].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += v; total += v
    print(f"# {path}: per-kernel device time (ncu gpu__time_duration.sum, cold-cache, serialised launches)")
    print(f"{'kernel':60s} {'launches':>9s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:60s} {n:9d} {t:12.1f} {t / n:10.2f} {100 * t / total:6.1f}%")
    print(f"{'TOTAL':60s} {sum(a[0] for a in agg.values()):9d} {total:12.1f}")

def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}: ncu --set full, one block per distinct kernel and grid (first captured launch of each)")
    seen = set()
    for r in rows[2:]:
        key = (r[idx["Kernel Name"]], r[idx["launch__grid_size"]] if "launch__grid_size" in idx else "")
        if key in seen:
            continue
        seen.add(key)
        print("----", r[idx["Kernel Name"]].split("(")[0], "grid", r[idx.get("launch__grid_size", 0)] if "launch__grid_size" in idx else "")
        for k in KEYS:
            if k in idx:
                print(f"  {k:85s} {r[idx[k]]:>16s} {units[idx[k]]}")

if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
