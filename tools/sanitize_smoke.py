#!/usr/bin/env python3
"""Every device entry point once on tiny inputs, meant to run under `compute-sanitizer --tool memcheck` (no oracle, no
timing): convert (plain, sharded, heavy-split forced), MC, DC, indexed MC, STL codec, ActiveVoxelsMesher, merge_points,
builders, CSG incl. active tiles, offset, the pipelined remesh, the two-step extraction, an open mesh."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import baby_shark_b200 as bs  # noqa: E402
from baby_shark_b200 import synth  # noqa: E402

L, ctx = bs.load_library(), bs.Context.default()
tris, vs, _ = synth.config_mesh(5, 0.03)
vol = bs.MeshToVolume().with_voxel_size(vs).convert(tris)
mc = bs.MarchingCubesMesher().with_voxel_size(vs).mesh(vol)
dc = bs.DualContouringMesher().with_voxel_size(vs).mesh(vol)
idx = bs.mesh_indexed(vol, vs)
assert np.array_equal(idx.points[idx.indices], mc)
stl = bs.StlWriter().write_to_buffer(mc)
back = bs.StlReader().read_from_buffer(stl)
assert back.n_tris == mc.shape[0] // 3
vol2 = bs.MeshToVolume().with_voxel_size(vs).convert(back)
boxes = bs.ActiveVoxelsMesher().mesh(vol)
os.environ["BSHARK_HEAVY_SHIFT"] = "5"
d_tris = torch.from_numpy(tris).cuda()
parts = []
for r in range(2):
    h = C.c_void_p()
    ctx.check(L.bs_mesh_to_volume_sharded(ctx._h, C.c_void_p(d_tris.data_ptr()), tris.shape[0], vs, 0, r, 2, C.byref(h)))
    v = bs.Volume(h, ctx)
    parts.append(bs.MarchingCubesMesher().with_voxel_size(vs).mesh(v))
del os.environ["BSHARK_HEAVY_SHIFT"]
assert np.array_equal(np.concatenate(parts), mc)
b = bs.VolumeBuilder().with_voxel_size(0.05)
u = b.sphere(0.6, (1.5, 0.3, 0.2)).union(b.sphere(2.0, (0.1, 0.2, 0.3)))   # leaves active tiles
mu = bs.MarchingCubesMesher().with_voxel_size(0.05).mesh(u)
s = b.cuboid((0, 0, 0), (1, 1, 1)).subtract(b.sphere(0.4, (0.9, 0.9, 0.9)))
i = b.iwp((0, 0, 0), (1, 1, 1), 0.5).intersect(b.sphere(0.5, (0.5, 0.5, 0.5)))
o1 = bs.MeshToVolume().with_voxel_size(0.1).convert(synth.cube()).offset(0.25)
o2 = bs.MeshToVolume().with_voxel_size(0.1).convert(synth.cube()).offset(-0.15)
# round 2: the one-call pipelined remesh (slabs, second stream, c-vertex carry), the two-step extraction into two destinations,
# the per-voxel sign path of an open mesh
rm = bs.VoxelRemesher().with_voxel_size(vs)
assert np.array_equal(rm.remesh(tris, 3), mc)
assert np.array_equal(rm.with_meshing_method(bs.MeshingMethod.FeaturePreserving).remesh(tris, 2), dc)
nv = C.c_size_t()
ctx.check(L.bs_mesh_mc_count(vol._h, vs, C.byref(nv)))
bufs = [torch.empty(nv.value * 3 + 64, dtype=torch.float32, device="cuda") for _ in range(2)]
ctx.check(L.bs_mesh_mc_emit_push(vol._h, (C.c_void_p * 2)(*[t.data_ptr() for t in bufs]), 2, 32, nv.value * 3 + 64))
torch.cuda.synchronize()
assert all(np.array_equal(t[32:32 + nv.value * 3].cpu().numpy().reshape(-1, 3), mc) for t in bufs)
cz = tris.reshape(-1, 3, 3)[:, :, 2].mean(1)
open_vol = bs.MeshToVolume().with_voxel_size(vs).convert(np.ascontiguousarray(tris[cz < np.median(cz)]))
assert ctx.last_stats()["sign_propagation"] == 0.0
print("sanitize smoke ok", mc.shape, dc.shape, idx.points.shape, boxes.shape, mu.shape, s.counts(), i.counts(), o1.counts(), o2.counts())
