#!/usr/bin/env python3
"""Debug aid: convert a BASELINE config on the GPU and dump the packed sign bits of the volume (plus origins / masks) to
an .npz that can be compared offline with the oracle's dump (tests/golden/make_config_hashes.py --dump).
usage: dump_signs.py CFG SCALE OUT.npz [0|1 sign propagation]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import baby_shark_b200 as bs  # noqa: E402
from baby_shark_b200 import synth, verify  # noqa: E402

cfg, scale, out = int(sys.argv[1]), float(sys.argv[2]), sys.argv[3]
prop = int(sys.argv[4]) if len(sys.argv) > 4 else 1
tris, vs, desc = synth.config_mesh(cfg, scale)
ctx = bs.Context.default()
ctx.set_flag(bs.BS_FLAG_SIGN_PROPAGATION, prop)
v = bs.MeshToVolume().with_voxel_size(vs).convert(tris)
st = ctx.last_stats()
d = v.download()
act = verify.active_bits(d["masks"])
np.savez_compressed(out, origins=d["origins"], masks=d["masks"], neg=np.packbits(np.signbit(d["values"]) & act, axis=1))
fp = verify.fingerprint_volume(d)
verts = bs.MarchingCubesMesher().with_voxel_size(vs).mesh(v)
fp.update(verify.fingerprint_soup(verts))
print(desc, fp, {k: v for k, v in st.items() if not k.endswith("_ms")})
