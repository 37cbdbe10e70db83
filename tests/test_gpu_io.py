"""GPU parity tests (-m gpu) for the rows either side of the path: STL codec, ActiveVoxelsMesher,
merge_points -- through the C ABI, bit-exact against the CPU oracle."""
import os
import struct

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def make_stl(tris, header=b"made by tests"):
    t = np.asarray(tris, np.float32).reshape(-1, 9)
    rec = np.zeros((t.shape[0], 50), np.uint8)
    rec[:, 12:48] = t.view(np.uint8).reshape(-1, 36)
    rec[:, 0:12] = np.float32([0.25, -1.5, 3.0]).view(np.uint8)  # normals in the file are ignored by the reader
    rec[:, 48] = 7                                                  # so is the attribute
    return header.ljust(80, b"\0") + struct.pack("<I", t.shape[0]) + rec.tobytes()


@pytest.mark.parametrize("n", [0, 1, 2, 255, 256, 257, 1001])
def test_stl_decode_matches_oracle(bs, oracle, n):
    rng = np.random.default_rng(n)
    tris = rng.normal(size=(n, 9)).astype(np.float32)
    data = make_stl(tris)
    got = bs.StlReader().read_from_buffer(data).numpy()
    ref = oracle.stl_decode(data)
    assert got.shape == ref.shape == (n, 9)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert np.array_equal(got.view(np.uint32), tris.view(np.uint32))


def test_stl_decode_reference_asset_and_convert(bs, oracle, box):
    # assets/box.stl as shipped by the reference (copied byte for byte by tests/golden/make_golden.py)
    data = open(os.path.join(HERE, "golden", "box.stl"), "rb").read()
    dev = bs.StlReader().read_from_buffer(data)
    assert np.array_equal(dev.numpy(), box.reshape(-1, 9))
    # decoded triangles feed the conversion without leaving the device
    a = bs.MeshToVolume().with_voxel_size(0.1).convert(dev).download()
    b = bs.MeshToVolume().with_voxel_size(0.1).convert(box).download()
    assert np.array_equal(a["origins"], b["origins"]) and np.array_equal(a["masks"], b["masks"])


def test_stl_short_buffer_is_an_error(bs, oracle):
    data = make_stl(np.zeros((10, 9), np.float32))
    for cut in (10, 83, 84 + 49, len(data) - 1):
        assert oracle.stl_decode(data[:cut]) is None
        with pytest.raises(bs.BsharkError):
            bs.StlReader().read_from_buffer(data[:cut])
    assert bs.StlReader().read_from_buffer(data + b"trailing bytes are ignored").n_tris == 10


def test_stl_encode_matches_oracle_bitwise(bs, oracle):
    from baby_shark_b200 import synth
    tris, vs, _ = synth.config_mesh(5, 0.03)
    soup = tris.reshape(-1, 3).copy()
    soup[9:12] = soup[9]                      # a degenerate face: zero normal
    soup[30:33] = [[0, 0, 0], [1, 0, 0], [2, 0, 0]]  # collinear
    got = bs.StlWriter().write_to_buffer(soup)
    ref = oracle.stl_encode(soup)
    assert len(got) == len(ref) == 84 + 50 * (soup.shape[0] // 3)
    assert got == ref
    # round trip through the reader
    assert np.array_equal(bs.StlReader().read_from_buffer(got).numpy(), soup.reshape(-1, 9))


def test_remesh_to_stl_round_trip(bs, oracle):
    # STL bytes -> device triangles -> volume -> MC -> STL bytes, compared with the oracle end to end
    from baby_shark_b200 import synth
    tris, vs, _ = synth.config_mesh(3, 0.05)
    data = make_stl(tris)
    vol = bs.MeshToVolume().with_voxel_size(vs).convert(bs.StlReader().read_from_buffer(data))
    verts = bs.MarchingCubesMesher().with_voxel_size(vs).mesh(vol)
    ovol, _ = oracle.mesh_to_volume(oracle.stl_decode(data), vs, 0, 8)
    overts = oracle.marching_cubes(ovol, vs)
    assert bs.StlWriter().write_to_buffer(verts) == oracle.stl_encode(overts)


def test_active_voxels_mesher(bs, oracle):
    from baby_shark_b200 import synth
    for cfg, scale in ((3, 0.04), (5, 0.03)):
        tris, vs, _ = synth.config_mesh(cfg, scale)
        g = bs.MeshToVolume().with_voxel_size(vs).convert(tris)
        o, _ = oracle.mesh_to_volume(tris, vs, 0, 8)
        gv, ov = bs.ActiveVoxelsMesher().mesh(g), oracle.active_voxels(o)
        assert gv.shape == ov.shape and gv.shape[0] % 3 == 0 and gv.shape[0] > 0
        assert np.array_equal(gv, ov)
    # a solid block spanning several bricks: only the outer faces are emitted, 2 triangles each
    blk = bs.VolumeBuilder().with_voxel_size(1.0).cuboid((-5.0, -5.0, -5.0), (12.0, 12.0, 12.0))
    oblk = oracle.cuboid(1.0, (-5.0, -5.0, -5.0), (12.0, 12.0, 12.0))
    assert np.array_equal(bs.ActiveVoxelsMesher().mesh(blk), oracle.active_voxels(oblk))
    assert bs.ActiveVoxelsMesher().mesh(bs.Volume.with_voxel_size(1.0)).shape == (0, 3)


def test_merge_points_matches_oracle(bs, oracle):
    from baby_shark_b200 import synth
    tris, vs, _ = synth.config_mesh(5, 0.04)
    g = bs.MeshToVolume().with_voxel_size(vs).convert(tris)
    soup = bs.MarchingCubesMesher().with_voxel_size(vs).mesh(g)
    got = bs.merge_points(soup)
    uq, idx = oracle.merge_points(soup)
    assert np.array_equal(got.indices, idx)
    assert np.array_equal(got.points.view(np.uint32), uq.view(np.uint32))
    assert np.array_equal(got.points[got.indices].view(np.uint32) & 0x7FFFFFFF, soup.view(np.uint32) & 0x7FFFFFFF)
    # a closed MC surface: every vertex is shared, Euler characteristic of a sphere
    n_v, n_f = got.points.shape[0], soup.shape[0] // 3
    assert n_v - 3 * n_f // 2 + n_f == 2


def test_merge_points_edge_cases(bs, oracle):
    pts = np.float32([[0.0, 1, 2], [-0.0, 1, 2], [np.nan, 0, 0], [np.nan, 0, 0], [5, 5, 5], [0.0, 1, 2], [5, 5, 5], [np.inf, 0, 0], [np.inf, 0, 0]])
    got = bs.merge_points(pts)
    uq, idx = oracle.merge_points(pts)
    assert list(got.indices) == list(idx) == [0, 0, 1, 2, 3, 0, 3, 4, 4]   # +0 == -0; NaN never merges
    assert np.array_equal(got.points.view(np.uint32), uq.view(np.uint32))
    assert bs.merge_points(np.zeros((0, 3), np.float32)).points.shape == (0, 3)
    rng = np.random.default_rng(1)
    many = rng.integers(0, 50, size=(200000, 3)).astype(np.float32)   # heavy duplication: long probe chains, atomicMin races
    got, (uq, idx) = bs.merge_points(many), oracle.merge_points(many)
    assert np.array_equal(got.indices, idx) and np.array_equal(got.points, uq)


def test_mesh_indexed_equals_mc_then_merge_points(bs, oracle):
    from baby_shark_b200 import synth
    tris, vs, _ = synth.config_mesh(4, 0.06)
    g = bs.MeshToVolume().with_voxel_size(vs).convert(tris)
    soup = bs.MarchingCubesMesher().with_voxel_size(vs).mesh(g)
    got = bs.mesh_indexed(g, vs)
    uq, idx = oracle.merge_points(soup)
    assert np.array_equal(got.indices, idx) and np.array_equal(got.points.view(np.uint32), uq.view(np.uint32))
    assert np.array_equal(got.points[got.indices].view(np.uint32), soup.view(np.uint32))


def test_active_voxels_mesher_on_active_tiles(bs, oracle):
    # a union that leaves active 8^3 tiles (the big sphere's interior slots): tiles contribute their boundary voxels --
    # edge and corner voxels several times, as in the reference -- and count as active neighbours of brick voxels
    vs = 0.05
    g = bs.VolumeBuilder().with_voxel_size(vs).sphere(0.6, (1.5, 0.3, 0.2)).union(bs.VolumeBuilder().with_voxel_size(vs).sphere(2.0, (0.1, 0.2, 0.3)))
    o = oracle.sphere(vs, 0.6, (1.5, 0.3, 0.2)).union(oracle.sphere(vs, 2.0, (0.1, 0.2, 0.3)))
    assert o.download()["tile_sizes"].size > 0
    gv, ov = bs.ActiveVoxelsMesher().mesh(g), oracle.active_voxels(o)
    assert gv.shape == ov.shape and gv.shape[0] > 0
    assert np.array_equal(gv, ov)
