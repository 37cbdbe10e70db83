#!/usr/bin/env python3
"""Opcode mix of one kernel from `ncu -i rep --page source --csv --kernel-name regex:NAME` output (executed counts, samples)."""
import collections
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, samp, tot = collections.Counter(), collections.Counter(), 0
for r in rows[hi + 1:]:
    if r and r[0] == "Address":  # the export repeats the listing: count it once
        break
    if len(r) <= iE or not r[iE].isdigit():
        continue
    t = r[iS].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    n = int(r[iE]); ops[op] += n; samp[op] += int(r[iSamp] or 0); tot += n
print("total warp instructions %d, per unit %.1f" % (tot, tot / div))
for op, n in ops.most_common(32):
    print("%-10s %6.2f%%  per unit %9.1f   stall samples %5.2f%%" % (op, 100 * n / tot, n / div, 100 * samp[op] / max(1, sum(samp.values()))))
