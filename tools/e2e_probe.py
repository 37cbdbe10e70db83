#!/usr/bin/env python3
"""Time bs_voxel_remesh_into (pinned host triangles in, pinned host vertices out) for several slab counts on one GPU:
`python tools/e2e_probe.py [config] [scale] [slabs,...]`. A probe for choosing the default; the bench's e2e leg is the number."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import baby_shark_b200 as bs  # noqa: E402
from baby_shark_b200 import synth  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 5
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
slabs = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 4, 8]
tris, vs, desc = synth.config_mesh(cfg, scale)
L = bs.load_library()
ctx = bs.Context.default()
h_tris = torch.from_numpy(tris).pin_memory()
h_out = torch.empty(16, dtype=torch.float32).pin_memory()
for K in slabs:
    ts = []
    for it in range(6):
        nf = C.c_size_t()
        t0 = time.perf_counter()
        st = L.bs_voxel_remesh_into(ctx._h, C.c_void_p(h_tris.data_ptr()), tris.shape[0], vs, 0, K, C.c_void_p(h_out.data_ptr()), h_out.numel(), C.byref(nf))
        dt = time.perf_counter() - t0
        if st == 3 and nf.value > h_out.numel():
            h_out = torch.empty(int(nf.value * 1.05) + 1024, dtype=torch.float32).pin_memory()
            continue
        ctx.check(st)
        ts.append(dt * 1e3)
    stt = ctx.last_stats()
    print("slabs %d: %.2f ms (min %.2f) floats %d  " % (K, sum(ts[1:]) / max(1, len(ts) - 1), min(ts), nf.value), {k: round(v, 2) for k, v in stt.items() if k.endswith("_ms")})
