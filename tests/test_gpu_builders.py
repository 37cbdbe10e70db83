"""GPU parity tests (-m gpu): VolumeBuilder primitives and Volume::from_fn through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

from util import compare_soups, compare_volumes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("vs,radius,origin", [(0.2, 3.0, (8, 8, 8)), (0.05, 0.44, (0.5, 0.5, 0.5)), (0.13, 1.0, (-0.3, 0.2, -1.7))])
def test_sphere(bs, oracle, vs, radius, origin):
    g = bs.VolumeBuilder().with_voxel_size(vs).sphere(radius, origin)
    o = oracle.sphere(vs, radius, origin)
    compare_volumes(g.download(), o.download(), vs)
    compare_soups(bs.MarchingCubesMesher().with_voxel_size(vs).mesh(g), oracle.marching_cubes(o, vs), vs, ordered=True)


@pytest.mark.parametrize("vs,mn,mx", [(0.2, (0, 0, 0), (10, 10, 10)), (0.07, (-1.03, -0.5, 0.2), (0.4, 0.77, 1.9))])
def test_cuboid(bs, oracle, vs, mn, mx):
    g = bs.VolumeBuilder().with_voxel_size(vs).cuboid(mn, mx)
    o = oracle.cuboid(vs, mn, mx)
    compare_volumes(g.download(), o.download(), vs)
    compare_soups(bs.MarchingCubesMesher().with_voxel_size(vs).mesh(g), oracle.marching_cubes(o, vs), vs, ordered=True)


def test_iwp_within_tolerance(bs, oracle):
    # device cosf vs host libm: values within a few ulp; the kept set can differ only where |f| is at the band limit
    vs = 0.1
    g = bs.VolumeBuilder().with_voxel_size(vs).iwp((0, 0, 0), (6, 6, 6), 1.0).download()
    o = oracle.iwp(vs, (0, 0, 0), (6, 6, 6), 1.0).download()
    from util import active_mask_bits
    assert np.array_equal(g["origins"], o["origins"])
    ga, oa = active_mask_bits(g["masks"]), active_mask_bits(o["masks"])
    both = ga & oa
    assert (ga != oa).sum() <= 1e-4 * oa.sum()
    assert np.abs(g["values"][both] - o["values"][both]).max() <= 1e-5 * vs


def test_from_fn_matches_oracle(bs, oracle):
    vs = 0.25
    f = lambda p: (np.linalg.norm(p - np.float32(1.0), axis=1) - np.float32(2.0)).astype(np.float32)  # noqa: E731
    g = bs.Volume.from_fn(vs, (-2, -2, -2), (4, 4, 4), 1, f)
    lo, hi = np.floor(np.float32(-2) / np.float32(vs)).astype(int), np.ceil(np.float32(4) / np.float32(vs)).astype(int)
    ijk = np.stack(np.meshgrid(*[np.arange(lo, hi + 1)] * 3, indexing="ij"), -1).reshape(-1, 3)
    val = f(ijk.astype(np.float32) * np.float32(vs))
    keep = ~(np.abs(val) > np.float32(2) * np.float32(vs))
    o = oracle.from_voxels(ijk[keep], val[keep], vs)
    compare_volumes(g.download(), o.download(), vs)


def test_from_voxels_duplicates_last_wins(bs, oracle):
    import ctypes as C
    ijk = np.array([[0, 0, 0], [1, 2, 3], [0, 0, 0], [-9, 4, 100], [1, 2, 3], [0, 0, 0]], np.int32)
    val = np.array([1, 2, 3, 4, 5, 6], np.float32)
    ctx = bs.Context.default()
    h = C.c_void_p()
    ctx.check(bs.load_library().bs_volume_from_voxels(ctx._h, ijk.ctypes.data_as(C.POINTER(C.c_int32)), val.ctypes.data_as(C.POINTER(C.c_float)), 6, 1.0, C.byref(h)))
    g = bs.Volume(h, ctx)
    compare_volumes(g.download(), oracle.from_voxels(ijk, val, 1.0).download(), 1.0)


def test_failed_calls_do_not_disturb_live_volumes(bs, oracle):
    # error paths return early; what they still hold goes back to the block cache when the next call starts -- and
    # nothing else: a volume created before the failing calls must stay intact
    import ctypes as C
    from baby_shark_b200 import synth
    tris, vs, _ = synth.config_mesh(3, 0.04)
    keep = bs.MeshToVolume().with_voxel_size(vs).convert(tris)
    before = keep.download()
    L, ctx = bs.load_library(), bs.Context.default()
    for _ in range(3):
        with pytest.raises(bs.BsharkError):
            bs.StlReader().read_from_buffer(b"\0" * 80 + (1000).to_bytes(4, "little") + b"\0" * 50)   # short buffer
        with pytest.raises(bs.BsharkError):
            bs.MeshToVolume().with_voxel_size(-1.0).convert(tris)                                        # invalid argument
        u = bs.VolumeBuilder().with_voxel_size(0.2).sphere(0.6, (1.5, 0.3, 0.2)).union(bs.VolumeBuilder().with_voxel_size(0.2).sphere(4.0, (0.1, 0.2, 0.3)))
        with pytest.raises(bs.ReferencePanic):
            bs.DualContouringMesher().with_voxel_size(0.2).mesh(u)                                      # todo!() on active tiles
        assert ctx.check(L.bs_context_copy_out_verts(ctx._h, None, 0)) is None
        other = bs.MeshToVolume().with_voxel_size(vs).convert(tris)                                     # reuses reclaimed blocks
        after = keep.download()
        assert np.array_equal(before["values"].view(np.uint32), after["values"].view(np.uint32)) and np.array_equal(before["masks"], after["masks"])
        assert np.array_equal(other.download()["values"].view(np.uint32), before["values"].view(np.uint32))
