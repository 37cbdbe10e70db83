#!/usr/bin/env python3
"""Golden fingerprints of the benchmark configs from the CPU oracle (run in the build container; minutes at scale 1).

  python tests/golden/make_config_hashes.py 5 1.0 [--threads 8] [--dump /tmp/cfg5.npz]

runs oracle convert (+ MC) on synth.config_mesh(cfg, scale) and stores the fingerprints (baby_shark_b200/verify.py) under
key "cfg<cfg>@<scale>" of tests/golden/config_hashes.json. bench.py and the slow GPU tests assert them on the device
output. --dump also writes the sign bits / origins for debugging a mismatch (not committed)."""
import argparse, json, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from baby_shark_b200 import synth, verify  # noqa: E402
from oracle import oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("cfg", type=int)
ap.add_argument("scale", type=float)
ap.add_argument("--threads", type=int, default=os.cpu_count())
ap.add_argument("--dump", default=None)
ap.add_argument("--no-mc", action="store_true")
a = ap.parse_args()
if a.cfg not in (1, 3, 4, 5):
    raise SystemExit("convert + MC configs: 1, 3, 4, 5")
tris, vs, desc = synth.config_mesh(a.cfg, a.scale)
t0 = time.time()
vol, st = O.mesh_to_volume(tris, vs, 0, a.threads)
t1 = time.time()
d = vol.download()
fp = verify.fingerprint_volume(d)
fp.update({"desc": desc, "n_tris": int(tris.shape[0]), "voxel_size": vs, "oracle_convert_s": round(t1 - t0, 1),
           "oracle_stage_s": {k: round(st[k], 2) for k in ("t_subdivide", "t_tree", "t_udf", "t_sign")}, "threads": a.threads})
if not a.no_mc:
    t2 = time.time()
    verts = O.marching_cubes(vol, vs)
    fp.update(verify.fingerprint_soup(verts))
    fp["oracle_mc_s"] = round(time.time() - t2, 1)
if a.dump:
    act = verify.active_bits(d["masks"])
    np.savez_compressed(a.dump, origins=d["origins"], masks=d["masks"], neg=np.packbits(np.signbit(d["values"]) & act, axis=1))
path = os.path.join(ROOT, "tests", "golden", "config_hashes.json")
allh = json.load(open(path)) if os.path.exists(path) else {}
allh["cfg%d@%g" % (a.cfg, a.scale)] = fp
json.dump(allh, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(fp))
