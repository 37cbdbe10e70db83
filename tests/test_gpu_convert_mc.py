"""GPU parity tests (-m gpu): mesh -> volume and marching cubes through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

from baby_shark_b200 import synth
from util import compare_soups, compare_volumes

pytestmark = pytest.mark.gpu


def convert_both(bs, oracle, tris, vs, band=0):
    g = bs.MeshToVolume().with_voxel_size(vs).with_narrow_band_width(band).convert(tris)
    o, st = oracle.mesh_to_volume(tris, vs, band, threads=8)
    return g, o, st


def test_box2_known_answer_topology(bs, oracle, box2):
    # survey intermediates of volume/mod.rs:134-152: 1854 active voxels in 8 leaves, 980 negative
    g, o, st = convert_both(bs, oracle, box2, 0.2)
    c = g.counts()
    assert (c["leaves"], c["active"], c["negative"]) == (8, 1854, 980)
    compare_volumes(g.download(), o.download(), 0.2)


@pytest.mark.parametrize("vs", [0.2, 0.1, 0.05, 0.031])
def test_box_grid_aligned(bs, oracle, box, vs):
    # box.stl is grid aligned: one ulp in the sequential subdivision sums flips floor/ceil (SURVEY hard part 1)
    g, o, _ = convert_both(bs, oracle, box, vs)
    compare_volumes(g.download(), o.download(), vs)


@pytest.mark.parametrize("band", [0, 1, 3])
def test_sphere_bands(bs, oracle, band):
    from baby_shark_b200 import synth
    tris = synth.uv_sphere(48, 24, 0.4, (0.503, 0.504, 0.505))
    g, o, _ = convert_both(bs, oracle, tris, 1.0 / 48, band)
    compare_volumes(g.download(), o.download(), 1.0 / 48)


def test_negative_coordinates_and_root_boundaries(bs, oracle):
    # a mesh straddling the origin spans 8 root nodes of the reference's tree; brick order must still match
    from baby_shark_b200 import synth
    tris = synth.uv_sphere(32, 16, 1.0, (0.01, -0.02, 0.03))
    g, o, _ = convert_both(bs, oracle, tris, 0.05)
    compare_volumes(g.download(), o.download(), 0.05)


def test_ragged_inputs(bs, oracle):
    # degenerate (zero-area) triangles, a sliver, a triangle smaller than a voxel, a large one, duplicates
    tris = np.array([
        [0, 0, 0, 0, 0, 0, 0, 0, 0],
        [0.1, 0.1, 0.1, 0.9, 0.1, 0.1, 0.5, 0.1, 0.1],
        [0.2, 0.2, 0.2, 0.21, 0.2, 0.2, 0.2, 0.21, 0.2],
        [-1.0, -1.0, 0.3, 1.5, -1.0, 0.35, 0.2, 1.7, 0.4],
        [-1.0, -1.0, 0.3, 1.5, -1.0, 0.35, 0.2, 1.7, 0.4],
        [0.0, 0.0, 1.0, 1.0, 0.0, 1.0, 1.0, 1e-7, 1.0],
    ], np.float32)
    g, o, _ = convert_both(bs, oracle, tris, 0.07)
    compare_volumes(g.download(), o.download(), 0.07, sign_tie=np.inf)  # open soup: only topology and |d| are defined


def test_empty_mesh_returns_none(bs):
    assert bs.MeshToVolume().with_voxel_size(0.1).convert(np.zeros((0, 9), np.float32)) is None


def test_mc_box2_matches_oracle_in_order(bs, oracle, box2):
    g, o, _ = convert_both(bs, oracle, box2, 0.2)
    gv = bs.MarchingCubesMesher().with_voxel_size(0.2).mesh(g)
    ov = oracle.marching_cubes(o, 0.2)
    compare_soups(gv, ov, 0.2, ordered=True)


def test_mc_sphere_and_noise_sphere(bs, oracle):
    from baby_shark_b200 import synth
    for cfg, scale in ((3, 0.06), (4, 0.08), (5, 0.04)):
        tris, vs, _ = synth.config_mesh(cfg, scale)
        g, o, _ = convert_both(bs, oracle, tris, vs)
        compare_volumes(g.download(), o.download(), vs)
        gv = bs.MarchingCubesMesher().with_voxel_size(vs).mesh(g)
        ov = oracle.marching_cubes(o, vs)
        assert gv.shape[0] > 1000
        compare_soups(gv, ov, vs, ordered=True)


def test_mc_ambiguous_cases_random_field(bs, oracle):
    # random signed values exercise every MC33 case incl. the face / interior tests and the c-vertex
    # (the reference has no tests for these branches; the oracle is the pin)
    rng = np.random.default_rng(0)
    n = 20
    ijk = np.stack(np.meshgrid(*[np.arange(-3, n - 3)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.int32)
    val = rng.uniform(-1, 1, ijk.shape[0]).astype(np.float32)
    val[rng.random(ijk.shape[0]) < 0.02] = 0.0
    val[rng.random(ijk.shape[0]) < 0.02] = -0.0
    keep = rng.random(ijk.shape[0]) < 0.97
    try:
        g = bs.Volume.from_fn  # noqa: F841  (from_voxels path)
        import ctypes as C
        L = bs.load_library()
        ctx = bs.Context.default()
        h = C.c_void_p()
        a, b = np.ascontiguousarray(ijk[keep]), np.ascontiguousarray(val[keep])
        st = L.bs_volume_from_voxels(ctx._h, a.ctypes.data_as(C.POINTER(C.c_int32)), b.ctypes.data_as(C.POINTER(C.c_float)), a.shape[0], 0.5, C.byref(h))
        if st == bs.BS_ERR_UNSUPPORTED:
            pytest.skip("bs_volume_from_voxels not implemented yet")
        ctx.check(st)
        gvol = bs.Volume(h, ctx)
    finally:
        pass
    ovol = oracle.from_voxels(ijk[keep], val[keep], 0.5)
    compare_volumes(gvol.download(), ovol.download(), 0.5)
    gv = bs.MarchingCubesMesher().with_voxel_size(0.5).mesh(gvol)
    ov, st = oracle.marching_cubes(ovol, 0.5, with_stats=True)
    assert all(st.case_hist[c] > 0 for c in range(1, 15)), list(st.case_hist)
    # includes tiling 6.1.2, where the reference emits the c-vertex left behind by an earlier cell
    # (marching_cubes.rs:103-111 never calls compute_c_vertex for it): the device carries that value through the
    # brick scan, so even those triangles are bit-identical
    compare_soups(gv, ov, 0.5, ordered=True)


def test_voxel_remesher_cube(bs):
    # src/remeshing/voxel.rs:105-112
    from baby_shark_b200 import synth
    v = bs.VoxelRemesher().with_voxel_size(0.1).remesh(synth.cube())
    assert v is not None and v.shape[0] > 0 and v.shape[0] % 3 == 0


@pytest.mark.parametrize("world", [2, 3])
def test_brick_sharded_outputs_concatenate_to_single_gpu_output(bs, world):
    # bs_mesh_to_volume_sharded: rank r keeps slab r (+ halo); MC / DC emit owned bricks only; concatenating the
    # per-rank outputs in rank order must reproduce the unsharded output exactly (run here on one device)
    import ctypes as C
    import torch
    from baby_shark_b200 import synth
    tris, vs, _ = synth.config_mesh(5, 0.04)
    L, ctx = bs.load_library(), bs.Context.default()
    d_tris = torch.from_numpy(tris).cuda()
    full = bs.MeshToVolume().with_voxel_size(vs).convert(tris)
    mc_full = bs.MarchingCubesMesher().with_voxel_size(vs).mesh(full)
    dc_full = bs.DualContouringMesher().with_voxel_size(vs).mesh(full)
    mc_parts, dc_parts, owned = [], [], 0
    for r in range(world):
        h = C.c_void_p()
        ctx.check(L.bs_mesh_to_volume_sharded(ctx._h, C.c_void_p(d_tris.data_ptr()), tris.shape[0], vs, 0, r, world, C.byref(h)))
        owned += ctx.last_stats()["n_bricks_owned"]
        v = bs.Volume(h, ctx)
        mc_parts.append(bs.MarchingCubesMesher().with_voxel_size(vs).mesh(v))
        dc_parts.append(bs.DualContouringMesher().with_voxel_size(vs).mesh(v))
    assert owned == full.counts()["leaves"]
    compare_soups(np.concatenate(mc_parts), mc_full, vs, ordered=True)
    compare_soups(np.concatenate(dc_parts), dc_full, vs, ordered=True)


def test_bunny_sign_agreement(bs, oracle, bunny):
    # config 1 mesh (assets/bunny.stl: 13 000 triangles, closed, genus 0) at a coarser voxel so the CPU oracle finishes in seconds.
    # Topology and |d| must be bit-exact. Signs come from different trees (device LBVH + hoisting vs the reference's
    # SAH tree): they could differ only where the winding number is close to the 0.2 threshold; on this closed mesh none do.
    vs = 0.5
    g = bs.MeshToVolume().with_voxel_size(vs).convert(bunny).download()
    o = oracle.mesh_to_volume(bunny, vs, 0, 8)[0].download()
    from util import active_mask_bits
    assert np.array_equal(g["origins"], o["origins"])
    m = active_mask_bits(o["masks"])
    assert np.array_equal(active_mask_bits(g["masks"]), m)
    gv, ov = g["values"][m], o["values"][m]
    assert np.array_equal(np.abs(gv).view(np.uint32), np.abs(ov).view(np.uint32))
    diff = np.signbit(gv) != np.signbit(ov)
    frac = diff.mean()
    print("bunny: %d active voxels, %d sign disagreements (%.4f%%)" % (gv.size, int(diff.sum()), 100 * frac))
    assert frac < 2e-3
    if diff.any():
        # every disagreement sits where the exact winding number is near the threshold
        idx = np.argwhere(m)[diff]
        pts = (o["origins"][idx[:, 0]] + np.stack([idx[:, 1] >> 6, (idx[:, 1] >> 3) & 7, idx[:, 1] & 7], 1)).astype(np.float32) * np.float32(vs)
        wn_exact, _ = oracle.winding_numbers(bunny, pts[:200], beta=-1.0)
        assert (np.abs(wn_exact - 0.2) < 0.15).all(), wn_exact


@pytest.mark.parametrize("shift", [4, 7])
def test_heavy_brick_split_gives_the_same_volume(bs, oracle, shift, monkeypatch):
    # bricks under many triangles are signed by 16 warps per work item, each walking a share of an expanded root list
    # (bs_fwn.cu "heavy bricks"); BSHARK_HEAVY_SHIFT lowers the threshold so that small test meshes take that path
    from baby_shark_b200 import synth
    monkeypatch.setenv("BSHARK_HEAVY_SHIFT", str(shift))
    monkeypatch.setenv("BSHARK_NO_BRUTE", "1")  # the split belongs to the LBVH traversal: keep closed meshes on it
    for cfg, scale in ((5, 0.05), (3, 0.06)):
        tris, vs, _ = synth.config_mesh(cfg, scale)
        g = bs.MeshToVolume().with_voxel_size(vs).convert(tris)
        assert bs.Context.default().last_stats()["n_heavy_bricks"] > 0
        o, _ = oracle.mesh_to_volume(tris, vs, 0, 8)
        compare_volumes(g.download(), o.download(), vs)
    monkeypatch.delenv("BSHARK_HEAVY_SHIFT")
    g2 = bs.MeshToVolume().with_voxel_size(vs).convert(tris)
    assert np.array_equal(g.download()["values"].view(np.uint32), g2.download()["values"].view(np.uint32))


def test_heavy_split_on_bunny_matches_unsplit_signs(bs, bunny, monkeypatch):
    monkeypatch.setenv("BSHARK_NO_BRUTE", "1")
    a = bs.MeshToVolume().with_voxel_size(0.5).convert(bunny).download()
    monkeypatch.setenv("BSHARK_HEAVY_SHIFT", "5")
    b = bs.MeshToVolume().with_voxel_size(0.5).convert(bunny).download()
    assert bs.Context.default().last_stats()["n_heavy_bricks"] > 0
    from util import active_mask_bits
    m = active_mask_bits(a["masks"])
    diff = np.signbit(a["values"][m]) != np.signbit(b["values"][m])
    assert diff.mean() < 2e-3  # the split walk refines some far nodes: only voxels at the 0.2 threshold may move


@pytest.mark.parametrize("world", [2, 4, 7])
def test_sharded_ranks_partition_the_volume_exactly(bs, world):
    # every rank derives the same coarse cut and works on its own triangle list; owned bricks must partition the unsharded
    # brick list in order, with identical values, and every rank's halo must carry exact values too
    import ctypes as C
    import torch
    from baby_shark_b200 import synth
    from util import active_mask_bits
    tris, vs, _ = synth.config_mesh(5, 0.125)
    L, ctx = bs.load_library(), bs.Context.default()
    d_tris = torch.from_numpy(tris).cuda()
    full = bs.MeshToVolume().with_voxel_size(vs).convert(tris).download()
    index = {tuple(o): i for i, o in enumerate(full["origins"].tolist())}
    fa = active_mask_bits(full["masks"])
    owned_total, nonempty = 0, 0
    for r in range(world):
        h = C.c_void_p()
        ctx.check(L.bs_mesh_to_volume_sharded(ctx._h, C.c_void_p(d_tris.data_ptr()), tris.shape[0], vs, 0, r, world, C.byref(h)))
        st = ctx.last_stats()
        owned_total += int(st["n_bricks_owned"])
        nonempty += st["n_bricks_owned"] > 0
        assert st["n_tris_local"] < tris.shape[0] or world == 1
        d = bs.Volume(h, ctx).download()
        if d["origins"].shape[0] == 0:
            continue
        idx = np.array([index[tuple(o)] for o in d["origins"].tolist()])  # every kept brick exists in the unsharded volume
        assert np.all(np.diff(idx) > 0)
        a = active_mask_bits(d["masks"])
        assert np.array_equal(a, fa[idx])
        assert np.array_equal(d["values"][a].view(np.uint32), full["values"][idx][a].view(np.uint32))
    assert owned_total == full["origins"].shape[0]
    assert nonempty >= min(world, 2)


@pytest.mark.parametrize("slabs", [1, 2, 3, 8])
def test_pipelined_remesh_equals_convert_then_mesh(bs, slabs):
    # VoxelRemesher::remesh as one call (bs_voxel_remesh_into): slab-wise convert + extraction with the read-back of one
    # slab overlapping the next -- the same vertices, bit for bit and in order, as convert followed by mesh
    tris, vs, _ = synth.config_mesh(5, 0.06)
    vol = bs.MeshToVolume().with_voxel_size(vs).convert(tris)
    for method, mesher in ((bs.MeshingMethod.Manifold, bs.MarchingCubesMesher()), (bs.MeshingMethod.FeaturePreserving, bs.DualContouringMesher())):
        ref = mesher.with_voxel_size(vs).mesh(vol)
        r = bs.VoxelRemesher().with_voxel_size(vs).with_meshing_method(method)
        out = np.full(ref.size + 64, np.nan, np.float32)
        n = r.remesh_into(np.ascontiguousarray(tris, np.float32), out, slabs)
        assert n == ref.size
        assert np.array_equal(out[:n].view(np.uint32), ref.reshape(-1).view(np.uint32))
        assert np.isnan(out[n:]).all()
        if slabs > 1:
            assert bs.Context.default().last_stats()["remesh_slabs"] == slabs
    # a buffer that is too small reports the size to retry with; the public wrapper retries by itself
    small = np.empty(1000, np.float32)
    need = bs.VoxelRemesher().with_voxel_size(vs).remesh_into(np.ascontiguousarray(tris, np.float32), small, slabs)
    assert need > small.size
    got = bs.VoxelRemesher().with_voxel_size(vs).remesh(tris, slabs)
    mc = bs.MarchingCubesMesher().with_voxel_size(vs).mesh(vol)
    assert np.array_equal(got.view(np.uint32), mc.view(np.uint32))
    assert bs.VoxelRemesher().with_voxel_size(vs).remesh(np.zeros((0, 9), np.float32)) is None


def test_two_step_marching_cubes_emits_into_every_destination(bs):
    # bs_mesh_mc_count + bs_mesh_mc_emit_push (the multi-GPU output exchange: the emit kernel stores every triangle into all
    # ranks' buffers): here both destinations live on this GPU; each must hold the one-call result at the given offset
    import ctypes as C
    L, ctx = bs.load_library(), bs.Context.default()
    tris, vs, _ = synth.config_mesh(5, 0.06)
    vol = bs.MeshToVolume().with_voxel_size(vs).convert(tris)
    ref = bs.MarchingCubesMesher().with_voxel_size(vs).mesh(vol).reshape(-1)
    nv = C.c_size_t()
    ctx.check(L.bs_mesh_mc_count(vol._h, vs, C.byref(nv)))
    assert nv.value * 3 == ref.size
    off, cap = 100, ref.size + 164
    ptrs = []
    for _ in range(2):
        p, handle = C.c_void_p(), C.create_string_buffer(64)
        ctx.check(L.bs_ipc_alloc(ctx._h, cap * 4, C.byref(p), handle))
        ptrs.append(p.value)
    arr = (C.c_void_p * 2)(*ptrs)
    try:
        ctx.check(L.bs_mesh_mc_emit_push(vol._h, arr, 2, off, cap))
        for p in ptrs:
            out = np.empty(ref.size, np.float32)
            ctx.check(L.bs_copy_to_host(ctx._h, C.c_void_p(p + 4 * off), C.c_void_p(out.ctypes.data), out.nbytes))
            assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
        assert L.bs_mesh_mc_emit_push(vol._h, arr, 2, off, cap) == 3  # the count was consumed: BS_ERR_INVALID
        ctx.check(L.bs_mesh_mc_count(vol._h, vs, C.byref(nv)))
        assert L.bs_mesh_mc_emit_push(vol._h, arr, 2, off, ref.size) == 3  # destination too small
    finally:
        for p in ptrs:
            L.bs_ipc_free(ctx._h, C.c_void_p(p))


def test_pipelined_remesh_with_empty_slabs(bs):
    # 12 triangles cut into 8 slabs: most slabs own nothing
    tris = synth.cube((0.1, -0.2, 0.3), 1.0, 0.8, 0.6)
    vs = 0.05
    ref = bs.MarchingCubesMesher().with_voxel_size(vs).mesh(bs.MeshToVolume().with_voxel_size(vs).convert(tris))
    got = bs.VoxelRemesher().with_voxel_size(vs).remesh(tris, 8)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # a mesh that produces no voxel at all is None, however many slabs it is cut into (mesh_to_volume.rs:58-60)
    nan = np.full((5, 9), np.nan, np.float32)
    assert bs.VoxelRemesher().with_voxel_size(vs).remesh(nan, 1) is None
    assert bs.VoxelRemesher().with_voxel_size(vs).remesh(nan, 4) is None
