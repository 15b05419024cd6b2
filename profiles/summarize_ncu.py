#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics plus the kernel split into regions of equal execution count (loop bodies)
with their share of executed instructions and of stall samples.
Usage: summarize_ncu.py file.ncu-rep [kernel-name-regex]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
flt = ["-k", "regex:" + sys.argv[2]] if len(sys.argv) > 2 else []
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + flt, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "inst_executed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("== kernel:", r[h.index("Kernel Name")][:60])
    for n in want:
        if n in h:
            print("  %-62s %s %s" % (n, r[h.index(n)], rows[1][h.index(n)]))
    for i, n in enumerate(h):
        if "issue_stalled" in n and n.endswith("per_warp_active.pct") and float(r[i] or 0) > 3:
            print("  stall %-56s %s" % (n.replace("smsp__warp_issue_stalled_", "").replace("_per_warp_active.pct", ""), r[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + flt, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
h = rows[hi]
si, ei, ti, ss = h.index("Source"), h.index("Instructions Executed"), h.index("Avg. Threads Executed"), h.index("# Samples")
data = [(r[si].strip(), int(r[ei]), float(r[ti]), int(r[ss])) for r in rows[hi + 1:] if len(r) > ei and r[ei].isdigit()]
tot = sum(d[1] for d in data)
stot = sum(d[3] for d in data)
print("SASS instructions %d, executed %d, samples %d" % (len(data), tot, stot))
seg, cur = [], None
for i, (s, e, t, sm) in enumerate(data):
    if cur is None or abs(e - cur[0]) > 0.02 * max(e, cur[0]):
        cur = [e, i, i, 0, 0]
        seg.append(cur)
    cur[2] = i
    cur[3] += e
    cur[4] += sm
for lvl, a, b, e, sm in seg:
    if e > 0.01 * tot or sm > 0.01 * stot:
        print("  sass[%4d-%4d] n=%4d exec/instr=%9d  instr %5.1f%%  samples %5.1f%%  threads %4.1f  %s"
              % (a, b, b - a + 1, lvl, 100.0 * e / tot, 100.0 * sm / stot, data[a][2], data[a][0][:44]))
