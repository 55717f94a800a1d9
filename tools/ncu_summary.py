#!/usr/bin/env python3
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

  python tools/ncu_summary.py launches <launches.csv> <out.txt>     # per-kernel totals and shares
  python tools/ncu_summary.py kernel <prof.ncu-rep> <out.txt>       # key counters of every captured launch
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "launch__occupancy_limit_registers", "sm__maximum_warps_per_active_cycle_pct", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_thread_inst_executed_op_integer_pred_on.sum"]


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki])
        v = float(r[vi].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[ui], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# {path}: {len(data)} launches, {tot / 1e6:.3f} ms total (ncu: cold-cache, serialised - compare SHARES)\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{a[1] / 1e6:10.3f} ms {a[0]:5d}x {100 * a[1] / tot:5.1f}%  avg {a[1] / a[0] / 1e3:10.1f} us  {k[:110]}\n")
    print(open(out).read())


def kernel(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out, "w") as f:
        f.write(f"# {path}: ncu --set full, {len(data)} launch(es)\n")
        ki = hdr.index("Kernel Name")
        for r in data:
            f.write(f"\n== {r[ki][:140]}\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"  {k:85s} {r[i]:>18s} {units[i]}\n")
    print(open(out).read())


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
