"""GPU parity tests (-m gpu): flood fill + CSG through the C ABI vs the CPU oracle (volumes incl. active tiles,
then marching cubes of the result in the reference's emission order)."""
import numpy as np
import pytest

from util import compare_soups, compare_volumes

pytestmark = pytest.mark.gpu
OPS = ["union", "subtract", "intersect"]


def both_prims(bs, oracle, vs):
    gb, ob = bs.VolumeBuilder().with_voxel_size(vs), oracle
    return (gb.cuboid((0, 0, 0), (10, 10, 10)), gb.sphere(3.0, (8, 8, 8))), (ob.cuboid(vs, (0, 0, 0), (10, 10, 10)), ob.sphere(vs, 3.0, (8, 8, 8)))


@pytest.mark.parametrize("op", OPS)
def test_cuboid_sphere(bs, oracle, op):
    # examples/dual_contouring.rs:11-18 shapes, all three operations
    vs = 0.2
    (ga, gb_), (oa, ob) = both_prims(bs, oracle, vs)
    g = getattr(ga, op)(gb_)
    o = getattr(oa, op)(ob)
    rep = compare_volumes(g.download(), o.download(), vs)
    assert rep["bricks"] > 0
    compare_soups(bs.MarchingCubesMesher().with_voxel_size(vs).mesh(g), oracle.marching_cubes(o, vs), vs, ordered=True)


@pytest.mark.parametrize("op", OPS)
@pytest.mark.parametrize("swap", [False, True])
def test_big_small_spheres_create_tiles(bs, oracle, op, swap):
    # a small sphere inside / straddling a big one: union makes whole 8^3 slots "inside" -> active -MAX tiles
    vs = 0.05
    def mk(B, big):
        return (B.sphere(2.0, (0.1, 0.2, 0.3)) if big else B.sphere(0.6, (1.5, 0.3, 0.2))) if B is not oracle else \
            (oracle.sphere(vs, 2.0, (0.1, 0.2, 0.3)) if big else oracle.sphere(vs, 0.6, (1.5, 0.3, 0.2)))
    gb = bs.VolumeBuilder().with_voxel_size(vs)
    ga, gb_ = mk(gb, not swap), mk(gb, swap)
    oa, ob = mk(oracle, not swap), mk(oracle, swap)
    g, o = getattr(ga, op)(gb_), getattr(oa, op)(ob)
    gd, od = g.download(), o.download()
    compare_volumes(gd, od, vs)
    if op == "union" and swap:  # B = the big sphere: its interior slots that A does not cover become active tiles
        assert od["tile_sizes"].size > 0, "this case is meant to exercise active tiles"
    compare_soups(bs.MarchingCubesMesher().with_voxel_size(vs).mesh(g), oracle.marching_cubes(o, vs), vs, ordered=True)


@pytest.mark.parametrize("op", ["union", "subtract"])
def test_two_tori_config2_scaled(bs, oracle, op):
    from baby_shark_b200 import synth
    (ta, tb), vs, _ = synth.config_mesh(2, 0.125)
    ga, gb_ = [bs.MeshToVolume().with_voxel_size(vs).convert(t) for t in (ta, tb)]
    oa, ob = [oracle.mesh_to_volume(t, vs, 0, 8)[0] for t in (ta, tb)]
    g, o = getattr(ga, op)(gb_), getattr(oa, op)(ob)
    compare_volumes(g.download(), o.download(), vs)
    compare_soups(bs.MarchingCubesMesher().with_voxel_size(vs).mesh(g), oracle.marching_cubes(o, vs), vs, ordered=True)


def test_boolean_example_chain(bs, oracle):
    # examples/boolean.rs:21-29 in miniature: union a row of boxes, then intersect / union / subtract with a sphere
    vs = 0.25
    gb = bs.VolumeBuilder().with_voxel_size(vs)
    gboxes, oboxes = bs.Volume.with_voxel_size(vs), oracle.empty(vs)
    for x in range(-8, 9, 3):
        gboxes = gboxes.union(gb.cuboid((x, -6.0, 0.0), (x + 1.0, 6.0, 12.0)))
        oboxes = oboxes.union(oracle.cuboid(vs, (x, -6.0, 0.0), (x + 1.0, 6.0, 12.0)))
    compare_volumes(gboxes.download(), oboxes.download(), vs)
    for op in OPS:
        g = getattr(gb.sphere(7.0, (0.0, 0.0, 6.0)), op)(gboxes.clone())
        o = getattr(oracle.sphere(vs, 7.0, (0.0, 0.0, 6.0)), op)(oboxes.clone())
        compare_volumes(g.download(), o.download(), vs)
        compare_soups(bs.MarchingCubesMesher().with_voxel_size(vs).mesh(g), oracle.marching_cubes(o, vs), vs, ordered=True)


def test_union_with_self_and_empty(bs, oracle):
    vs = 0.1
    gb = bs.VolumeBuilder().with_voxel_size(vs)
    a = gb.sphere(1.0, (0.03, 0.02, 0.01))
    before = a.clone().download()
    u = a.clone().union(a.clone()).download()
    from util import active_mask_bits
    assert np.array_equal(before["origins"], u["origins"])
    m = active_mask_bits(before["masks"])
    assert np.array_equal(m, active_mask_bits(u["masks"])) and np.array_equal(before["values"][m], u["values"][m])
    e = a.clone().union(bs.Volume.with_voxel_size(vs)).download()
    assert np.array_equal(before["origins"], e["origins"]) and np.array_equal(before["values"][m], e["values"][m])
    assert a.clone().intersect(bs.Volume.with_voxel_size(vs)).counts()["leaves"] == 0


@pytest.mark.parametrize("op", OPS)
@pytest.mark.parametrize("swap", [False, True])
def test_root_level_flood_fill_fills_the_gap_between_inside_nodes(bs, oracle, op, swap):
    # root_node/flood_fill.rs:17-41: two 4096^3 root nodes on one z-line, both "inside" at their facing ends and not
    # adjacent, get the keys between them filled with empty all-negative nodes; a sphere sitting in that gap is then
    # (partly) swallowed by a union, turned into 32768 active tiles when it is the left operand, etc.
    ijk = np.int32([(x, y, z) for z0 in (4000, 8300) for x in range(10, 13) for y in range(10, 13) for z in range(z0, z0 + 3)])
    val = np.full(ijk.shape[0], -0.5, np.float32)
    def blobs(B):
        return oracle.from_voxels(ijk, val, 1.0) if B is oracle else bs.Volume.from_voxels(ijk, val, 1.0)
    def ball(B):
        return oracle.sphere(1.0, 20.0, (11.0, 11.0, 6000.0)) if B is oracle else bs.VolumeBuilder().with_voxel_size(1.0).sphere(20.0, (11.0, 11.0, 6000.0))
    ga, gb_ = (ball(bs), blobs(bs)) if swap else (blobs(bs), ball(bs))
    oa, ob = (ball(oracle), blobs(oracle)) if swap else (blobs(oracle), ball(oracle))
    g, o = getattr(ga, op)(gb_), getattr(oa, op)(ob)
    gd, od = g.download(), o.download()
    compare_volumes(gd, od, 1.0)
    if op == "union":
        assert od["origins"].shape[0] == 68                      # part of the ball is gone: the gap node is "inside"
        assert (od["tile_sizes"].size == 32768) == swap            # make_child_inside on every slot of the gap node
