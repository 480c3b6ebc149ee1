#!/usr/bin/env python
"""Operand-delivery model (see rf_model.py) over the STATIC SASS of a kernel's main loop.

rf_model.py needs an ncu source export (executed counts per instruction).  This script applies
the same cost rule to `cuobjdump -sass` output, so that a change can be judged in the build
container before any GPU time is spent:

  * the main loop = the span of the LAST backward branch of the function;
  * a predicated forward branch is assumed TAKEN when the span it skips contains a CALL (the
    rare repair paths) and NOT taken otherwise; uniform forward branches (BRA.U) are assumed
    taken (the multicast / peer-store alternatives of the extract kernel);
  * everything else inside the loop counts once per warp iteration.

    python profiles/static_rf.py <lib.so> <substring of the mangled kernel name> [-v]

Calibration against the ncu-based model of the committed capture (profiles/r2_rf_model.txt):
extract_blk_kernel<1,4> 1,423 cycles, embed_blk_kernel<3,1,true,false> 2,754.
"""
import collections
import re
import subprocess
import sys

NO_DEST = ('STG', 'STL', 'STS', 'BRA', 'BSYNC', 'BSSY', 'EXIT', 'RET', 'LDGSTS', 'CALL', 'WARPSYNC', 'NOP', 'BAR', 'ATOMS', 'RED')


def functions(lib):
    text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, out = None, {}
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = out.setdefault(m.group(1), [])
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur is not None:
            cur.append((int(m.group(1), 16), m.group(2).strip()))
    return out


def loop_body(ins):
    """[(addr, text)] executed once per iteration under the assumptions in the module docstring."""
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    back = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:!?U?P\w+,\s*)?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) <= a and 'WARPSYNC' not in t:
            # the out-of-line WARPSYNC stubs after EXIT also branch backwards: only count loops that
            # start before the first EXIT-terminated region, i.e. the largest span
            span = a - int(m.group(1), 16)
            if back is None or span > back[2]:
                back = (addr_index[int(m.group(1), 16)], i, span)
    lo, hi, _ = back
    body, i = [], lo
    while i <= hi:
        a, t = ins[i]
        m = re.match(r"^(@!?U?P\w+\s+)?BRA(\.\w+)*\s+(?:!?U?P\w+,\s*)?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(3), 16)
            uniform = '.U' in (m.group(2) or '') or (m.group(1) or '').startswith('@U') or (m.group(1) or '').startswith('@!U')
            div = '.DIV' in t
            if tgt > a and not div and tgt in addr_index and addr_index[tgt] <= hi + 1:
                skipped = ins[i + 1:addr_index[tgt]]
                rare = any(x[1].split()[0].startswith('CALL') or ' CALL' in x[1] for x in skipped)
                pred = m.group(1) is not None or re.search(r"BRA(\.\w+)*\s+!?U?P", t)
                if not pred or rare or '.U' in t:
                    body.append((a, t))
                    i = addr_index[tgt]
                    continue
        body.append((a, t))
        i += 1
    return body


def model(body, verbose=False):
    tot = fma = 0.0
    per_op, per_n = collections.Counter(), collections.Counter()
    prev_reuse = set()
    for a, src in body:
        m = re.match(r"^(@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)\s*(.*)$", src)
        if not m:
            prev_reuse = set()
            continue
        op, ops = m.group(2), m.group(3)
        opn = op.split('.')[0]
        parts = [p.strip() for p in ops.split(',')]
        srcs = parts if opn in NO_DEST else (parts[2:] if opn in ('ISETP', 'FSETP', 'PLOP3') else parts[1:])
        even = odd = 0
        reuse_next, seen = set(), set()
        for p in srcs:
            for mm in re.finditer(r"(?<![UP])R(\d+)((?:\.[A-Za-z0-9_]+)*)", p):
                rn, mods = int(mm.group(1)), mm.group(2)
                wide = 'F32x2' in mods or '.64' in mods
                regs = [rn, rn + 1] if wide else [rn]
                if 'reuse' in mods:
                    reuse_next.update(regs)
                for x in regs:
                    if x in prev_reuse or x in seen:
                        continue
                    seen.add(x)
                    if x % 2 == 0:
                        even += 1
                    else:
                        odd += 1
        prev_reuse = reuse_next
        clk = 0 if opn in ('BRA', 'BSYNC', 'BSSY', 'NOP') else max(even, odd, 1)
        tot += clk
        per_op[opn] += clk
        per_n[opn] += 1
        if opn in ('FADD2', 'FFMA2', 'FMUL2', 'IDP', 'IMAD'):
            fma += 2
        elif opn in ('FADD', 'FMUL', 'FFMA', 'FMNMX'):
            fma += 1
        if verbose:
            print("%05x %d %s" % (a, clk, src))
    return tot, fma, per_op, per_n


def main():
    lib, key = sys.argv[1], sys.argv[2]
    verbose = '-v' in sys.argv[3:]
    fns = functions(lib)
    names = [n for n in fns if key in n]
    for n in names:
        body = loop_body(fns[n])
        tot, fma, per_op, per_n = model(body, verbose)
        print("== %s\nstatic loop body: %d instructions, operand-delivery cycles %d, FP32-pipe cycles %d" % (n, sum(per_n.values()), tot, fma))
        for k, v in per_op.most_common(16):
            print("  %-8s n=%4d  rf-clk=%5d  (%.2f per instruction)" % (k, per_n[k], v, v / per_n[k]))


if __name__ == "__main__":
    main()
