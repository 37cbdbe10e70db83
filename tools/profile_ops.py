#!/usr/bin/env python3
"""One pass of the CSG / offset / dual-contouring rows on their BASELINE config (2 / 3 / 4): the target of ncu captures.
usage: profile_ops.py CFG [SCALE]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import baby_shark_b200 as bs  # noqa: E402
from baby_shark_b200 import synth  # noqa: E402
cfg = int(sys.argv[1]); scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
mesh, vs, desc = synth.config_mesh(cfg, scale)
ctx = bs.Context.default()
conv = lambda t: bs.MeshToVolume().with_voxel_size(vs).convert(t)  # noqa: E731
if cfg == 2:
    a, b = conv(mesh[0]), conv(mesh[1])
    a.clone().union(b.clone()); print(ctx.last_stats())
    a.subtract(b); print(ctx.last_stats())
elif cfg == 3:
    a = conv(mesh)
    a.clone().offset(2 * vs); print(ctx.last_stats())
    a.offset(-2 * vs); print(ctx.last_stats())
else:
    a = conv(mesh)
    bs.DualContouringMesher().with_voxel_size(vs).mesh(a); print(ctx.last_stats())
