#!/usr/bin/env python
"""Summarise an .ncu-rep into a small text file for profiles/ (details + selected raw metrics +
optionally the hottest source lines).  usage: ncu_summary.py <rep> <out.txt> "<title>" [cubin kernel-substr]"""
import csv
import io
import subprocess
import sys

KEEP = ("Duration", "Executed Ipc", "Issue Slots Busy", "No Eligible", "Eligible Warps", "Warp Cycles Per Issued",
        "Avg. Active Threads", "Achieved Occ", "Theoretical Occ", "Registers Per", "L1/TEX Hit", "L2 Hit",
        "DRAM Throughput", "Memory Throughput", "Dynamic Shared", "Static Shared", "Block Limit", "Compute (SM) Throughput")
RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
       "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum",
       "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size",
       "smsp__thread_inst_executed_per_inst_executed.ratio", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep, out, title = sys.argv[1:4]
    lines = ["# " + title, "# source: %s (ncu --set full --clock-control none)" % rep.split("/")[-1], ""]
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout.splitlines()
    for l in det:
        if any(k in l for k in KEEP) or ("(" in l and ")x(" in l):
            lines.append(l.rstrip())
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    lines += ["", "# raw metrics (per launch)"]
    if len(rows) > 2:
        h, units = rows[0], rows[1]
        for r in rows[2:]:
            for w in RAW:
                if w in h:
                    lines.append("%s = %s %s" % (w, r[h.index(w)], units[h.index(w)]))
            lines.append("")
    if len(sys.argv) > 5:
        lines += ["# hottest CUDA source lines (per-instruction counters joined with nvdisasm -g line info)"]
        import os
        tool = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_lines.py")
        lines += subprocess.run([sys.executable, tool, rep, sys.argv[4], sys.argv[5], "25"], capture_output=True,
                                text=True).stdout.splitlines()
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:12]))


if __name__ == "__main__":
    main()
