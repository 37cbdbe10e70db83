#!/usr/bin/env python3
"""Sum an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: name, launches, total ms, mean us.
usage: launch_times.py launches.csv [skip_first_n_launches]"""
import csv, sys, re, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
acc = collections.OrderedDict()
for r in rows[1 + skip:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v if r[ui] in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^.*::", "", name) if "cub" not in name else "cub::" + re.sub(r"<.*", "", name.split("::")[-1])
    a = acc.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in acc.values())
for k, a in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print("%-60s %5d launches %10.3f ms %8.1f us/launch %5.1f %%" % (k[:60], a[0], a[1] / 1e3, a[1] / a[0], 100 * a[1] / tot))
print("total %.3f ms" % (tot / 1e3))
