#!/usr/bin/env python
"""Dynamic SASS opcode mix from an `ncu --page source --csv` export: executed warp-instructions
per opcode, per warp (divide by the number of warps that entered the kernel)."""
import csv
import collections
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
mix, samples = collections.Counter(), collections.Counter()
total = 0
first = None
static = 0
for r in rows[2:]:
    if len(r) <= iex:
        continue
    src = r[isrc].strip()
    try:
        n = int(r[iex])
    except ValueError:
        continue
    if first is None:
        first = n
    static += 1
    op = re.sub(r"^@!?U?P\d+\s+", "", src).split()[0].split(".")[0]
    mix[op] += n
    samples[op] += int(r[ismp] or 0)
    total += n
print("static instructions: %d, executed warp-instr: %d, warps: %d, per warp: %.0f" % (static, total, first, total / first))
for op, n in mix.most_common(40):
    print("  %-10s %8.1f per warp   %5.1f%%   samples %d" % (op, n / first, 100.0 * n / total, samples[op]))
