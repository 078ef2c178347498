#!/usr/bin/env python
"""Attribute an ncu capture's per-instruction counters to CUDA source lines.

usage: ncu_lines.py <report.ncu-rep> <cubin> <mangled-kernel-substring> [top]
Joins `ncu --page source --csv` (SASS order, per-instruction counters) with `nvdisasm -g` line
markers of the same function in the cubin (needs -lineinfo at compile time).
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def main():
    rep, cubin, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    h = rows[hi]
    iS, iI, iN = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    iT = h.index("Thread Instructions Executed")
    sass = [(r[iS].strip(), int(r[iI]), int(r[iN]), int(r[iT])) for r in rows[hi + 1:] if len(r) > iT and r[iI].isdigit()]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    # locate the function
    start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l)
    cur = ("?", 0)
    lines = []
    for l in dis[start + 1:]:
        if l.startswith(".text.") or l.startswith("\t.section") and ".text." in l:
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            lines.append((cur, m.group(2)))
    if len(lines) != len(sass):
        print("warning: %d disasm instructions vs %d ncu rows" % (len(lines), len(sass)))
    agg = defaultdict(lambda: [0, 0, 0])
    n = min(len(lines), len(sass))
    for k in range(n):
        a = agg[lines[k][0]]
        a[0] += sass[k][1]; a[1] += sass[k][2]; a[2] += sass[k][3]
    ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
    print("total warp-instructions %d, samples %d, avg active threads %.1f" % (ti, ts, sum(v[2] for v in agg.values()) / max(ti, 1)))
    src_cache = {}
    for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        try:
            if f not in src_cache:
                import glob
                cand = glob.glob("/root/repo/auv-sim_b200/csrc/" + f)
                src_cache[f] = open(cand[0]).read().splitlines() if cand else []
            text = src_cache[f][ln - 1].strip()[:90] if src_cache[f] else ""
        except Exception:
            pass
        print("%5.2f%% inst %5.2f%% stall-samples %4.1f lanes  %-12s:%-4d %s" % (100.0 * v[0] / ti, 100.0 * v[1] / max(ts, 1),
                                                                             v[2] / max(v[0], 1), f, ln, text))


if __name__ == "__main__":
    main()
