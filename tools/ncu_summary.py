#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/launches_rNN.txt
    python tools/ncu_summary.py report   gpurun_out/prof.ncu-rep  > profiles/<kernel>_rNN.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fmaheavy.sum",
    "sm__inst_executed_pipe_fmalite.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__cycles_elapsed.avg.per_second",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = defaultdict(list)
    for r in rows[1:]:
        if len(r) > vi:
            agg[r[ki]].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    unit = rows[1][ui]
    print(f"# ncu --metrics gpu__time_duration.sum launch list ({path}); cold-cache, serialised: compare SHARES")
    print(f"# unit of the raw values: {unit}; total {tot / 1e6:.3f} ms over {sum(len(v) for v in agg.values())} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{sum(v) / tot:6.3f}  n={len(v):3d}  total={sum(v) / 1e6:9.3f} ms  avg={sum(v) / len(v) / 1e6:8.3f} ms  {k[:110]}")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"# {path}: {name}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"{w:90s} {vals[i]:>18s} {units[i]}")
        print()


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
