"""GPU parity tests (-m gpu): dual contouring through the C ABI vs the CPU oracle. The reference's output order is
nondeterministic across leaves (rayon + Mutex); both sides here emit leaves in visit order, so arrays compare 1:1."""
import numpy as np
import pytest

from util import compare_soups

pytestmark = pytest.mark.gpu


def test_dual_contouring_example(bs, oracle):
    # examples/dual_contouring.rs:11-18: cuboid.subtract(sphere) -> DC
    vs = 0.2
    gb = bs.VolumeBuilder().with_voxel_size(vs)
    g = gb.cuboid((0, 0, 0), (10, 10, 10)).subtract(gb.sphere(3.0, (8, 8, 8)))
    o = oracle.cuboid(vs, (0, 0, 0), (10, 10, 10)).subtract(oracle.sphere(vs, 3.0, (8, 8, 8)))
    gv = bs.DualContouringMesher().with_voxel_size(vs).mesh(g)
    ov = oracle.dual_contouring(o, vs)
    assert ov is not None and ov.shape[0] > 1000
    compare_soups(gv, ov, vs, ordered=True)


def test_dc_noise_sphere_config4_scaled(bs, oracle):
    from baby_shark_b200 import synth
    tris, vs, _ = synth.config_mesh(4, 0.08)
    g = bs.MeshToVolume().with_voxel_size(vs).convert(tris)
    o = oracle.mesh_to_volume(tris, vs, 0, 8)[0]
    gv = bs.DualContouringMesher().with_voxel_size(vs).mesh(g)
    ov = oracle.dual_contouring(o, vs)
    assert ov is not None and ov.shape[0] > 1000
    compare_soups(gv, ov, vs, ordered=True)


def test_dc_sharp_box_and_feature_preserving_remesher(bs, oracle, box2):
    g = bs.MeshToVolume().with_voxel_size(0.1).convert(box2)
    o = oracle.mesh_to_volume(box2, 0.1)[0]
    compare_soups(bs.DualContouringMesher().with_voxel_size(0.1).mesh(g), oracle.dual_contouring(o, 0.1), 0.1, ordered=True)
    v = bs.VoxelRemesher().with_voxel_size(0.1).with_meshing_method(bs.MeshingMethod.FeaturePreserving).remesh(box2)
    assert v.shape[0] > 0


def test_dc_panics_like_the_reference(bs, oracle):
    # active tiles -> todo!(); an isolated sign change without neighbours on an axis -> unreachable!()
    vs = 0.05
    gb = bs.VolumeBuilder().with_voxel_size(vs)
    g = gb.sphere(0.6, (1.5, 0.3, 0.2)).union(gb.sphere(2.0, (0.1, 0.2, 0.3)))
    with pytest.raises(bs.ReferencePanic):
        bs.DualContouringMesher().with_voxel_size(vs).mesh(g)
    import ctypes as C
    ijk = np.array([[0, 0, 0], [1, 0, 0]], np.int32)
    val = np.array([1.0, -1.0], np.float32)
    ctx = bs.Context.default()
    h = C.c_void_p()
    ctx.check(bs.load_library().bs_volume_from_voxels(ctx._h, ijk.ctypes.data_as(C.POINTER(C.c_int32)), val.ctypes.data_as(C.POINTER(C.c_float)), 2, 1.0, C.byref(h)))
    with pytest.raises(bs.ReferencePanic):
        bs.DualContouringMesher().mesh(bs.Volume(h, ctx))
    assert oracle.dual_contouring(oracle.from_voxels(ijk, val, 1.0)) is None
