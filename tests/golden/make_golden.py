#!/usr/bin/env python3
"""Generate the committed fixtures under tests/golden/ from the reference checkout (run in the build container
only: /root/reference does not exist on the GPU box).

  box2_tris.npy, box_tris.npy   the 12 triangles of assets/box2.stl / assets/box.stl as [n,9] f32 in file order
                                (vertex order matters: it fixes the subdivision direction, mesh_to_volume.rs:90-91)
  box.stl                       assets/box.stl byte for byte (684 bytes): the STL reader's known input (io/stl.rs)
  bunny_tris.npz                assets/bunny.stl (13 000 triangles), compressed
  reference_known_answers.json  the known answers the reference's own tests hold for this path plus the
                                survey-derived intermediates, with their source lines
"""
import json, os, struct, sys
import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def read_binary_stl(path):  # io/stl.rs:32-95: 80 B header, u32 count, 50 B records, normals ignored
    b = open(path, "rb").read()
    n = struct.unpack("<I", b[80:84])[0]
    rec = np.frombuffer(b[84:84 + 50 * n], dtype=np.uint8).reshape(n, 50)
    return rec[:, 12:48].copy().view("<f4").reshape(n, 9)


np.save(os.path.join(OUT, "box2_tris.npy"), read_binary_stl(os.path.join(REF, "assets/box2.stl")))
np.save(os.path.join(OUT, "box_tris.npy"), read_binary_stl(os.path.join(REF, "assets/box.stl")))
open(os.path.join(OUT, "box.stl"), "wb").write(open(os.path.join(REF, "assets/box.stl"), "rb").read())
np.savez_compressed(os.path.join(OUT, "bunny_tris.npz"), tris=read_binary_stl(os.path.join(REF, "assets/bunny.stl")))
json.dump({
    "test_volume_offset": {"source": "src/voxel/volume/mod.rs:134-152", "mesh": "box2_tris.npy", "voxel_size": 0.2,
                           "offset": 0.5, "mc_vertices": 7944},
    "survey_intermediates": {"source": "SURVEY.md section 4", "n_sub": 2352, "convert_active": 1854, "convert_leaves": 8,
                             "convert_negative": 980, "offset_active": 5348, "offset_leaves": 32,
                             "mc_case_hist": {"0": 2896, "1": 80, "2": 360, "5": 24, "8": 840, "9": 24}},
    "test_voxel_remeshing": {"source": "src/remeshing/voxel.rs:105-112", "mesh": "cube(1,1,1) @ origin", "voxel_size": 0.1,
                             "faces_gt": 0},
}, open(os.path.join(OUT, "reference_known_answers.json"), "w"), indent=1)
print("ok")
