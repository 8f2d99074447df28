#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md / bench.py cite.

usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_name.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instruction (warp execution efficiency x32)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % (dram__)"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("l1tex__t_bytes.sum", "L1 bytes"),
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALL_NAMES = ["long_scoreboard", "short_scoreboard", "wait", "not_selected", "math_pipe_throttle", "mio_throttle",
               "lg_throttle", "branch_resolving", "dispatch_stall", "no_instruction", "barrier", "selected"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu summary of `{rep}`\n")
    print("Captured with `ncu --set full --clock-control none --import-source on` (kernel replay; cold caches,")
    print("serialised: use shares and ratios, not absolute times).\n")
    for r in data:
        print(f"## {r[col['Kernel Name']][:110]}  (launch id {r[col['ID']]})\n")
        print("| metric | value |\n|---|---|")
        for k, label in KEYS:
            if k in col:
                print(f"| {label} (`{k}`) | {r[col[k]]} {units[col[k]]} |")
        print("\n| warp stall reason (warps per issue-active cycle) | value |\n|---|---|")
        for s in STALL_NAMES:
            k = STALLS % s
            if k in col:
                print(f"| {s} | {r[col[k]]} |")
        print()


if __name__ == "__main__":
    main()
