import sys,json
# usage: benchsum.py FILE   (or JSON lines on stdin when no file is given)
for l in (open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin):
    if l.startswith("{"):
        d=json.loads(l); print("ms/step", round(d["ms_per_step"],2), "value %.3g"%d["value"], {k:round(v,2) for k,v in d["stage_ms"].items()})
        print({k:v for k,v in d["work"].items() if not k.endswith("_ms")})
        print([(r["kernel"][:10], round(r["frac"],3), r.get("algorithmic")) for r in d["rooflines"]]); print(d["e2e"], d.get("clocks"))
        if "cpu_baseline" in d: print(d["cpu_baseline"])
    else: print(l, end="")
