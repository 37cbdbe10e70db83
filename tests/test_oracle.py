"""CPU tests (-m "not gpu"): the oracle against the reference's own known answers and golden fixtures."""
import json
import os

import numpy as np

from conftest import GOLDEN

KA = json.load(open(os.path.join(GOLDEN, "reference_known_answers.json")))


def test_reference_unit_tests_restated(oracle):
    # leaf CSG known answers, leaf/internal/root flood fill, tree insert/remove/fill (see bs_oracle.cpp:bso_selftest)
    assert oracle.selftest() == 0


def test_volume_offset_known_answer(oracle, box2):
    # src/voxel/volume/mod.rs:134-152 : box2.stl @0.2 -> offset(0.5) -> MC -> 7944 vertices
    ka, mid = KA["test_volume_offset"], KA["survey_intermediates"]
    vol, st = oracle.mesh_to_volume(box2, ka["voxel_size"])
    assert st["n_sub"] == mid["n_sub"]
    c = vol.counts()
    assert (c["active"], c["leaves"], c["negative"]) == (mid["convert_active"], mid["convert_leaves"], mid["convert_negative"])
    vol.offset(ka["offset"])
    c = vol.counts()
    assert (c["active"], c["leaves"]) == (mid["offset_active"], mid["offset_leaves"])
    verts, stats = oracle.marching_cubes(vol, with_stats=True)
    assert verts.shape[0] == ka["mc_vertices"]
    hist = {str(i): int(n) for i, n in enumerate(stats.case_hist) if n}
    assert hist == mid["mc_case_hist"]


def test_voxel_remeshing_cube(oracle):
    # src/remeshing/voxel.rs:105-112 : unit cube @0.1 -> faces > 0
    from baby_shark_b200 import synth
    vol, _ = oracle.mesh_to_volume(synth.cube(), 0.1)
    assert oracle.marching_cubes(vol).shape[0] > 0


def test_empty_mesh_is_none(oracle):
    vol, _ = oracle.mesh_to_volume(np.zeros((0, 9), np.float32), 0.1)
    assert vol is None


def test_closed_mesh_is_watertight(oracle):
    # invariant: a closed input gives a closed MC output -- every edge is shared by exactly 2 triangles
    from baby_shark_b200 import synth
    tris = synth.uv_sphere(24, 12, 0.4, (0.53, 0.54, 0.55))
    vol, _ = oracle.mesh_to_volume(tris, 1.0 / 32)
    v = oracle.marching_cubes(vol)
    t = v.reshape(-1, 3, 3)
    uniq, inv = np.unique(v.view(np.uint32).reshape(-1, 3), axis=0, return_inverse=True)
    f = inv.reshape(-1, 3)
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    e.sort(axis=1)
    _, counts = np.unique(e, axis=0, return_counts=True)
    assert t.shape[0] > 100 and (counts == 2).all()


def test_union_with_self_is_identity_on_band(oracle):
    from baby_shark_b200 import synth
    a, _ = oracle.mesh_to_volume(synth.uv_sphere(24, 12, 0.4, (0.5, 0.5, 0.5)), 1.0 / 32)
    before = a.clone().download()
    u = a.clone().union(a.clone()).download()
    from util import active_mask_bits
    ba, ua = active_mask_bits(before["masks"]), active_mask_bits(u["masks"])
    assert np.array_equal(before["origins"], u["origins"]) and np.array_equal(ba, ua)
    assert np.array_equal(before["values"][ba], u["values"][ua])


def test_builders_and_csg_dc(oracle):
    # examples/dual_contouring.rs:11-18 at a coarser voxel: cuboid.subtract(sphere) -> DC gives a mesh
    cube = oracle.cuboid(0.5, (0, 0, 0), (10, 10, 10))
    sph = oracle.sphere(0.5, 3.0, (8, 8, 8))
    v = oracle.dual_contouring(cube.subtract(sph))
    assert v is not None and v.shape[0] > 0 and v.shape[0] % 3 == 0


def test_offset_roundtrip_in_band(oracle):
    # offset(+d) then offset(-d) returns the original surface to within a fraction of a voxel (SURVEY 8c invariant)
    from baby_shark_b200 import synth
    vs = 1.0 / 32
    a, _ = oracle.mesh_to_volume(synth.uv_sphere(32, 16, 0.3, (0.5, 0.5, 0.5)), vs)
    v0 = oracle.marching_cubes(a.clone())
    v1 = oracle.marching_cubes(a.offset(2 * vs).offset(-2 * vs))
    r0 = np.linalg.norm(v0 - 0.5, axis=1).mean()
    r1 = np.linalg.norm(v1 - 0.5, axis=1).mean()
    assert abs(r0 - r1) < 0.5 * vs


def test_stl_codec_and_reference_asset(oracle, box):
    # io/stl.rs: the reader ignores header / normals / attributes; assets/box.stl decodes to the golden triangles
    import os
    data = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "box.stl"), "rb").read()
    t = oracle.stl_decode(data)
    assert np.array_equal(t, box.reshape(-1, 9))
    assert oracle.stl_decode(data[:-1]) is None
    out = oracle.stl_encode(t.reshape(-1, 3))
    assert len(out) == len(data) and out[:80] == b"\0" * 80 and out[80:84] == data[80:84]
    assert np.array_equal(oracle.stl_decode(out), t)
    # the writer recomputes unit normals (box.stl stores the same ones)
    n_out = np.frombuffer(out, np.uint8)[84:].reshape(-1, 50)[:, :12].copy().view(np.float32)
    n_in = np.frombuffer(data, np.uint8)[84:].reshape(-1, 50)[:, :12].copy().view(np.float32)
    assert np.allclose(np.linalg.norm(n_out, axis=1), 1.0, atol=1e-6) and np.allclose(n_out, n_in, atol=1e-6)


def test_active_voxels_and_merge_points_invariants(oracle):
    # one voxel: 6 exposed faces x 2 triangles; a 2-voxel bar hides the shared face
    v1 = oracle.from_voxels(np.int32([[3, 4, 5]]), np.float32([0.1]), 1.0)
    a = oracle.active_voxels(v1)
    assert a.shape == (36, 3) and a.min(0).tolist() == [3, 4, 5] and a.max(0).tolist() == [4, 5, 6]
    assert a[:6].tolist() == [[3, 4, 6], [4, 5, 6], [3, 5, 6], [3, 4, 6], [4, 4, 6], [4, 5, 6]]  # the top face comes first (active_voxels.rs:45-56)
    v2 = oracle.from_voxels(np.int32([[7, 0, 0], [8, 0, 0]]), np.float32([0.1, 0.2]), 1.0)  # across a leaf boundary
    assert oracle.active_voxels(v2).shape == (60, 3)
    uq, idx = oracle.merge_points(a.astype(np.float32))
    assert uq.shape == (8, 3) and idx.max() == 7 and np.array_equal(uq[idx], a.astype(np.float32))
    assert idx[:3].tolist() == [0, 1, 2]  # first-occurrence order


def test_half_edge_closedness_rule():
    # the rule bs_signprop.cu applies before propagating signs (numpy restatement in tools/sign_propagation_probe.py):
    # closed and consistently oriented <=> every directed edge once, its reverse once, no repeated vertex in a triangle
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
    from sign_propagation_probe import mesh_is_closed
    from baby_shark_b200 import synth
    sphere = synth.uv_sphere(24, 12, 1.0, (0.1, 0.2, 0.3))
    assert mesh_is_closed(sphere) and mesh_is_closed(synth.cube()) and mesh_is_closed(synth.torus(16, 8, 1.0, 0.3, (0, 0, 0)))
    assert not mesh_is_closed(sphere[:-1])                                   # a hole: boundary edges
    flipped = sphere.copy(); flipped[5] = flipped[5].reshape(3, 3)[[0, 2, 1]].reshape(9)
    assert not mesh_is_closed(flipped)                                       # one triangle with the other orientation
    assert not mesh_is_closed(np.concatenate([sphere, sphere[:1]]))          # a duplicated triangle: the same directed edges twice
    degenerate = sphere.copy(); degenerate[3, 3:6] = degenerate[3, 0:3]
    assert not mesh_is_closed(degenerate)                                    # repeated vertex inside a triangle
    assert mesh_is_closed(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bunny_tris.npz"))["tris"])
