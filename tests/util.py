"""Parity helpers: compare a device volume / vertex soup with the oracle's (north_star tolerances)."""
import numpy as np


def active_mask_bits(masks):
    """[n,8] u64 -> [n,512] bool, bit (o & 63) of word (o >> 6)."""
    m = np.ascontiguousarray(masks, np.uint64)
    bits = np.unpackbits(m.view(np.uint8).reshape(-1, 8, 8), axis=-1, bitorder="little")
    return bits.reshape(-1, 512).astype(bool)


def compare_volumes(gpu, ref, voxel_size, *, exact_values=True, sign_tie=1e-6, value_tol=1e-5):
    """gpu / ref: dicts from Volume.download(). Checks, in this order:
    brick list identical and in the same (reference visit) order; active masks bit-identical; |value| bit-identical
    (exact_values) or within value_tol*voxel_size; signs identical except where |d| < sign_tie*voxel_size.
    Returns a dict of counts for reporting."""
    assert gpu["origins"].shape == ref["origins"].shape, (gpu["origins"].shape, ref["origins"].shape)
    assert np.array_equal(gpu["origins"], ref["origins"]), "brick topology / order differs"
    ga, ra = active_mask_bits(gpu["masks"]), active_mask_bits(ref["masks"])
    assert np.array_equal(ga, ra), "active voxel masks differ: %d voxels" % int((ga != ra).sum())
    gv, rv = gpu["values"][ga], ref["values"][ra]
    if exact_values:
        same = np.abs(gv).view(np.uint32) == np.abs(rv).view(np.uint32)
        assert same.all(), "|value| bits differ at %d of %d voxels, max abs diff %g" % (int((~same).sum()), same.size, float(np.abs(np.abs(gv) - np.abs(rv)).max()))
    else:
        d = np.abs(np.abs(gv) - np.abs(rv))
        assert (d <= value_tol * voxel_size).all(), "max |dSDF| = %g voxels" % float(d.max() / voxel_size)
    sign_diff = np.signbit(gv) != np.signbit(rv)
    off_surface = np.abs(rv) >= sign_tie * voxel_size
    n_bad = int((sign_diff & off_surface).sum())
    assert n_bad == 0, "%d sign disagreements away from the surface (of %d active)" % (n_bad, gv.size)
    assert np.array_equal(gpu["tile_origins"], ref["tile_origins"]) and np.array_equal(gpu["tile_sizes"], ref["tile_sizes"])
    assert np.array_equal(gpu["tile_values"], ref["tile_values"])
    return dict(bricks=int(gpu["origins"].shape[0]), active=int(gv.size), sign_ties=int((sign_diff & ~off_surface).sum()))


def canonical_triangles(verts):
    """[3n,3] soup -> [n,9] with each triangle rotated to start at its lexicographically smallest vertex and the
    triangle list sorted: an order-free, orientation-preserving canonical form."""
    t = np.ascontiguousarray(verts, np.float32).reshape(-1, 3, 3)
    if t.shape[0] == 0:
        return t.reshape(0, 9)
    keys = t.view(np.uint32).astype(np.uint64)
    k = (keys[..., 0] << np.uint64(42)) ^ (keys[..., 1] << np.uint64(21)) ^ keys[..., 2]  # cheap per-vertex key
    first = np.argmin(k, axis=1)
    idx = (first[:, None] + np.arange(3)[None, :]) % 3
    t = np.take_along_axis(t, idx[:, :, None], axis=1).reshape(-1, 9)
    order = np.lexsort(t.T[::-1])
    return t[order]


def compare_soups(gpu_verts, ref_verts, voxel_size, *, ordered=True, tol=1e-5):
    """ordered: identical vertex arrays in the same order (bit-exact). Otherwise compare canonical sorted forms
    within tol*voxel_size."""
    assert gpu_verts.shape == ref_verts.shape, "vertex counts differ: %s vs %s" % (gpu_verts.shape, ref_verts.shape)
    if ordered:
        same = gpu_verts.view(np.uint32) == ref_verts.view(np.uint32)
        assert same.all(), "%d of %d vertex coordinates differ, max %g voxels" % (int((~same).sum()), same.size, float(np.abs(gpu_verts - ref_verts).max() / voxel_size))
        return
    a, b = canonical_triangles(gpu_verts), canonical_triangles(ref_verts)
    assert np.abs(a - b).max() <= tol * voxel_size, "max vertex difference %g voxels" % float(np.abs(a - b).max() / voxel_size)
