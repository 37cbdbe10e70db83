import sys, os, numpy as np, ctypes as C
sys.path.insert(0, '.')
import baby_shark_b200 as B
from baby_shark_b200 import synth
from oracle import oracle as O
# accuracy probe: wn sign margins are not observable directly; compare signs vs oracle on several meshes and report
for cfg, scale in ((5, 0.06), (4, 0.1), (3, 0.1)):
    tris, vs, _ = synth.config_mesh(cfg, scale)
    g = B.MeshToVolume().with_voxel_size(vs).convert(tris).download()
    o = O.mesh_to_volume(tris, vs, 0, 16)[0].download()
    m = np.unpackbits(np.ascontiguousarray(o["masks"]).view(np.uint8).reshape(-1, 8, 8), axis=-1, bitorder="little").reshape(-1, 512).astype(bool)
    d = np.signbit(g["values"][m]) != np.signbit(o["values"][m])
    print("kappa", os.environ.get("BSHARK_KAPPA"), "cfg", cfg, "voxels", m.sum(), "sign diffs", int(d.sum()))
(ta, tb), vs, _ = synth.config_mesh(2, 0.2)
for t in (ta, tb):
    g = B.MeshToVolume().with_voxel_size(vs).convert(t).download(); o = O.mesh_to_volume(t, vs, 0, 16)[0].download()
    m = np.unpackbits(np.ascontiguousarray(o["masks"]).view(np.uint8).reshape(-1, 8, 8), axis=-1, bitorder="little").reshape(-1, 512).astype(bool)
    print("torus sign diffs", int((np.signbit(g["values"][m]) != np.signbit(o["values"][m])).sum()), "of", m.sum())
bun = np.load("tests/golden/bunny_tris.npz")["tris"]
g = B.MeshToVolume().with_voxel_size(0.4).convert(bun).download(); o = O.mesh_to_volume(bun, 0.4, 0, 16)[0].download()
m = np.unpackbits(np.ascontiguousarray(o["masks"]).view(np.uint8).reshape(-1, 8, 8), axis=-1, bitorder="little").reshape(-1, 512).astype(bool)
print("bunny sign diffs", int((np.signbit(g["values"][m]) != np.signbit(o["values"][m])).sum()), "of", m.sum())
