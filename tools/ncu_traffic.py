#!/usr/bin/env python3
"""Per-kernel DRAM traffic and duration from an `ncu --set full` report -> JSON (committed under profiles/, read by bench.py
for `roofline.traffic`): tools/ncu_traffic.py REPORT.ncu-rep "workload description" > profiles/rN_ncu_traffic_cfg5.json"""
import csv, json, re, subprocess, sys
rep, desc = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
unit = dict(zip(hdr, units))
SC = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3, "ns": 1e-6, "us": 1e-3, "ms": 1.0}
res = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = re.sub(r"\(.*", "", d.get("Kernel Name", "?"))
    name = re.sub(r"^.*::", "", name)
    def val(k):
        return float(d[k].replace(",", "")) * SC.get(unit[k], 1.0)
    e = res.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "ms": 0.0})
    e["launches"] += 1
    e["dram_bytes"] += val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    e["ms"] += val("gpu__time_duration.sum")
for e in res.values():
    e["dram_bytes_per_launch"] = e.pop("dram_bytes") / e["launches"]
    e["ms_per_launch"] = e.pop("ms") / e["launches"]
print(json.dumps({"source": rep.split("/")[-1], "workload": desc, "how": "ncu --set full --clock-control none (cold caches, serialised): dram__bytes_read.sum + dram__bytes_write.sum per launch", "kernels": res}, indent=1))
