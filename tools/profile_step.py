#!/usr/bin/env python3
"""One convert + marching-cubes pass of a BASELINE config through the C ABI, nothing else: the target of the ncu captures
(`ncu --set full -k regex:... python tools/profile_step.py 5 1.0`). Numbers printed under a profiler are not bench values."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import baby_shark_b200 as bs  # noqa: E402
from baby_shark_b200 import synth  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 5
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
tris, vs, desc = synth.config_mesh(cfg, scale)
L = bs.load_library(os.environ.get("BSHARK_LIB"))  # experiment builds: BSHARK_LIB=path/to/variant.so
ctx = bs.Context.default()
d_tris = torch.from_numpy(tris).cuda()
for _ in range(reps):
    h = C.c_void_p()
    ctx.check(L.bs_mesh_to_volume_device(ctx._h, C.c_void_p(d_tris.data_ptr()), tris.shape[0], vs, 0, C.byref(h)))
    st = ctx.last_stats()
    dv, nv = C.c_void_p(), C.c_size_t()
    ctx.check(L.bs_mesh_mc_device(h, vs, C.byref(dv), C.byref(nv)))
    st2 = ctx.last_stats()
    L.bs_volume_free(h)
print(desc, {k: round(v, 2) for k, v in st.items() if k.endswith("_ms")}, {k: round(v, 2) for k, v in st2.items() if k.endswith("_ms")}, nv.value)
