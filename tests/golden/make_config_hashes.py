#!/usr/bin/env python3
"""Golden fingerprints of the benchmark configs from the CPU oracle (run in the build container; minutes at scale 1).

  python tests/golden/make_config_hashes.py 5 1.0 [--threads 8] [--dump /tmp/cfg5.npz]

runs the oracle on synth.config_mesh(cfg, scale) -- convert + MC (configs 1, 3, 4, 5), plus the row the config is about:
union / subtract + MC (2), offset(+2 voxels) / offset(-2 voxels) + MC (3), dual contouring (4) -- and stores the
fingerprints (baby_shark_b200/verify.py) under key "cfg<cfg>@<scale>" of tests/golden/config_hashes.json. bench.py and
tests/test_gpu_fullsize.py assert them on the device output. --dump also writes the sign bits / origins of the converted
volume for debugging a mismatch (not committed)."""
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
a = ap.parse_args()
mesh, vs, desc = synth.config_mesh(a.cfg, a.scale)
out = {"desc": desc, "voxel_size": vs, "threads": a.threads}
t0 = time.time()


def convert(tris, tag=""):
    t = time.time()
    vol, st = O.mesh_to_volume(tris, vs, 0, a.threads)
    d = vol.download()
    out.update({tag + k: v for k, v in verify.fingerprint_volume(d).items()})
    out[tag + "n_tris"] = int(tris.shape[0])
    out[tag + "oracle_convert_s"] = round(time.time() - t, 1)
    out[tag + "oracle_stage_s"] = {k: round(st[k], 2) for k in ("t_subdivide", "t_tree", "t_udf", "t_sign")}
    return vol, d


def mc(vol, tag=""):
    t = time.time()
    out.update({tag + k: v for k, v in verify.fingerprint_soup(O.marching_cubes(vol, vs)).items()})
    out[tag + "oracle_mc_s"] = round(time.time() - t, 1)


if a.cfg == 2:
    va, _ = convert(mesh[0], "a_")
    vb, _ = convert(mesh[1], "b_")
    for op in ("union", "subtract"):
        r = getattr(va.clone(), op)(vb.clone())
        out.update({op + "_" + k: v for k, v in verify.fingerprint_volume(r.download()).items()})
        mc(r, op + "_")
else:
    vol, d = convert(mesh)
    if a.dump:
        act = verify.active_bits(d["masks"])
        np.savez_compressed(a.dump, origins=d["origins"], masks=d["masks"], neg=np.packbits(np.signbit(d["values"]) & act, axis=1))
    mc(vol)
    if a.cfg == 3:
        for sgn, name in ((2.0, "offset_plus_"), (-2.0, "offset_minus_")):
            t = time.time()
            r = vol.clone().offset(np.float32(sgn) * np.float32(vs))
            out.update({name + k: v for k, v in verify.fingerprint_volume(r.download()).items()})
            out[name + "oracle_s"] = round(time.time() - t, 1)
            mc(r, name)
    if a.cfg == 4:
        t = time.time()
        out.update({"dc_" + k: v for k, v in verify.fingerprint_soup(O.dual_contouring(vol, vs)).items()})
        out["oracle_dc_s"] = round(time.time() - t, 1)
out["oracle_total_s"] = round(time.time() - t0, 1)
path = os.path.join(ROOT, "tests", "golden", "config_hashes.json")
allh = json.load(open(path)) if os.path.exists(path) else {}
allh["cfg%d@%g" % (a.cfg, a.scale)] = out
json.dump(allh, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(out))
