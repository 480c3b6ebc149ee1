#!/usr/bin/env python
"""Per code region (consecutive SASS lines, `chunk` instructions each): executed instructions and
stall-sample breakdown from an `ncu --page source --csv` export."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 250
hdr = rows[1]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
names = ["stall_barrier", "stall_long_sb", "stall_math", "stall_no_inst", "stall_not_selected", "stall_selected",
         "stall_short_sb", "stall_wait", "stall_dispatch", "stall_lg", "stall_branch_resolving", "stall_mio"]
idx = [hdr.index(n) for n in names]
need = max([iex, ismp] + idx)
body = [r for r in rows[2:2 + (len(rows) - 2) // 2] if len(r) > need]
print("%-12s %10s %8s  " % ("lines", "executed", "samples") + " ".join("%7s" % n.replace("stall_", "")[:7] for n in names))
tot = [0] * len(names)
for s in range(0, len(body), chunk):
    part = body[s:s + chunk]
    ex = sum(int(r[iex] or 0) for r in part)
    smp = sum(int(r[ismp] or 0) for r in part)
    vals = [sum(int(r[i] or 0) for r in part) for i in idx]
    tot = [a + b for a, b in zip(tot, vals)]
    print("%5d-%-6d %10d %8d  " % (s, s + len(part), ex, smp) + " ".join("%7d" % v for v in vals))
print("%-12s %10s %8s  " % ("total", "", "") + " ".join("%7d" % v for v in tot))
