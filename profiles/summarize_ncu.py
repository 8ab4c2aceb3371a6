#!/usr/bin/env python3
"""Summarise an `ncu --set full` report (.ncu-rep) into the handful of numbers DESIGN.md / bench.py quote.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rNN_<kernel>.txt
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, unit = rows[0], rows[1]
    for row in rows[2:]:
        d = dict(zip(head, row))
        u = dict(zip(head, unit))
        print("kernel:", d.get("Kernel Name"))
        for k in KEYS:
            if k in d:
                print("  %-72s %s %s" % (k, d[k], u[k]))
        for k in head:
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                try:
                    if float(d[k]) >= 0.1:
                        print("  %-72s %s" % (k.replace("smsp__average_warps_issue_stalled_", "stall/issue: "), d[k]))
                except ValueError:
                    pass
        print()


if __name__ == "__main__":
    main()
