"""GPU parity tests (-m gpu): Volume::offset (prune + fast sweeping + shift) through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

from util import compare_soups, compare_volumes

pytestmark = pytest.mark.gpu


def test_reference_known_answer_7944(bs, oracle, box2):
    # src/voxel/volume/mod.rs:134-152: box2.stl @0.2 -> offset(0.5) -> MC -> 7944 vertices; 5348 voxels in 32 leaves
    g = bs.MeshToVolume().with_voxel_size(0.2).convert(box2).offset(0.5)
    c = g.counts()
    assert (c["active"], c["leaves"]) == (5348, 32)
    o = oracle.mesh_to_volume(box2, 0.2)[0].offset(0.5)
    compare_volumes(g.download(), o.download(), 0.2)
    gv = bs.MarchingCubesMesher().with_voxel_size(0.2).mesh(g)
    assert gv.shape[0] == 7944
    compare_soups(gv, oracle.marching_cubes(o, 0.2), 0.2, ordered=True)


@pytest.mark.parametrize("d_vox", [2.0, -2.0, 0.7, -3.5, 6.0])
def test_sphere_offsets(bs, oracle, d_vox):
    from baby_shark_b200 import synth
    vs = 1.0 / 48
    tris = synth.uv_sphere(48, 24, 0.3, (0.503, 0.504, 0.505))
    g = bs.MeshToVolume().with_voxel_size(vs).convert(tris).offset(d_vox * vs)
    o = oracle.mesh_to_volume(tris, vs, 0, 8)[0].offset(d_vox * vs)
    compare_volumes(g.download(), o.download(), vs)
    compare_soups(bs.MarchingCubesMesher().with_voxel_size(vs).mesh(g), oracle.marching_cubes(o, vs), vs, ordered=True)


def test_offset_config3_scaled_both_signs(bs, oracle):
    from baby_shark_b200 import synth
    tris, vs, _ = synth.config_mesh(3, 0.08)
    for d in (2 * vs, -2 * vs):
        g = bs.MeshToVolume().with_voxel_size(vs).convert(tris).offset(d)
        o = oracle.mesh_to_volume(tris, vs, 0, 8)[0].offset(d)
        compare_volumes(g.download(), o.download(), vs)


def test_offset_after_union_drops_tiles(bs, oracle):
    # active tiles (+-MAX) are removed by the prune (volume/mod.rs:96), so offset after a union is defined
    vs = 0.05
    gb = bs.VolumeBuilder().with_voxel_size(vs)
    g = gb.sphere(0.6, (1.5, 0.3, 0.2)).union(gb.sphere(2.0, (0.1, 0.2, 0.3))).offset(0.12)
    o = oracle.sphere(vs, 0.6, (1.5, 0.3, 0.2)).union(oracle.sphere(vs, 2.0, (0.1, 0.2, 0.3))).offset(0.12)
    compare_volumes(g.download(), o.download(), vs)


def test_offset_of_empty_volume(bs):
    assert bs.Volume.with_voxel_size(0.1).offset(0.3).counts()["leaves"] == 0
