"""GPU tests (-m gpu) of sign propagation on closed meshes (bs_signprop.cu): the default sign path for closed inputs
must give the volume the per-voxel winding-number path gives, and the oracle's; open inputs must fall back."""
import numpy as np
import pytest

from util import active_mask_bits, compare_volumes

pytestmark = pytest.mark.gpu


def convert(bs, tris, vs, prop, band=0):
    ctx = bs.Context.default()
    ctx.set_flag(bs.BS_FLAG_SIGN_PROPAGATION, 1 if prop else 0)
    try:
        v = bs.MeshToVolume().with_voxel_size(vs).with_narrow_band_width(band).convert(tris)
        return v, ctx.last_stats()
    finally:
        ctx.set_flag(bs.BS_FLAG_SIGN_PROPAGATION, 1)


def active_values(d):
    return d["values"][active_mask_bits(d["masks"])]


@pytest.mark.parametrize("cfg,scale", [(5, 0.05), (3, 0.06), (4, 0.08), (5, 0.125)])
def test_propagated_signs_equal_per_voxel_signs_and_the_oracle(bs, oracle, cfg, scale, monkeypatch):
    from baby_shark_b200 import synth
    tris, vs, _ = synth.config_mesh(cfg, scale)
    a, st_a = convert(bs, tris, vs, prop=False)
    b, st_b = convert(bs, tris, vs, prop=True)
    assert st_a["sign_propagation"] == 0.0 and st_b["sign_propagation"] == 1.0
    da, db = a.download(), b.download()
    assert np.array_equal(da["masks"], db["masks"])
    assert np.array_equal(active_values(da).view(np.uint32), active_values(db).view(np.uint32))
    # only a handful of voxels is evaluated: the shells' representatives and voxels on the surface within rounding
    assert st_b["n_sign_seeds"] < 1e-4 * st_b["n_active"] + 16, st_b
    assert st_b["sign_brute_force"] == 1.0
    o, _ = oracle.mesh_to_volume(tris, vs, 0, threads=8)
    compare_volumes(db, o.download(), vs)
    # the same representatives through the LBVH traversal instead of the brute-force sum
    monkeypatch.setenv("BSHARK_NO_BRUTE", "1")
    c, st_c = convert(bs, tris, vs, prop=True)
    assert st_c["sign_propagation"] == 1.0 and st_c["sign_brute_force"] == 0.0
    assert np.array_equal(active_values(c.download()).view(np.uint32), active_values(db).view(np.uint32))


def test_tori_and_bands(bs, oracle):
    from baby_shark_b200 import synth
    (ta, tb), vs, _ = synth.config_mesh(2, 0.1)
    for tris, band in ((ta, 0), (tb, 0), (ta, 2)):
        v, st = convert(bs, tris, vs, prop=True, band=band)
        assert st["sign_propagation"] == 1.0
        o, _ = oracle.mesh_to_volume(tris, vs, band, threads=8)
        compare_volumes(v.download(), o.download(), vs)


def test_bunny_propagated_signs_match_the_oracle(bs, oracle, bunny):
    # closed genus-0 mesh with large triangles relative to the voxel (subdivision n ~ 2..3 at vs = 0.5): the
    # running-sum drift term of the tolerance is exercised
    vs = 0.5
    v, st = convert(bs, bunny, vs, prop=True)
    assert st["sign_propagation"] == 1.0
    o, _ = oracle.mesh_to_volume(bunny, vs, 0, threads=8)
    compare_volumes(v.download(), o.download(), vs)


def test_open_mesh_falls_back_to_per_voxel_signs(bs):
    from baby_shark_b200 import synth
    tris, vs, _ = synth.config_mesh(3, 0.06)
    _, st = convert(bs, tris[:-5], vs, prop=True)  # a few triangles missing: boundary edges
    assert st["sign_propagation"] == 0.0
    flipped = tris.copy()
    flipped[7] = flipped[7].reshape(3, 3)[[0, 2, 1]].reshape(9)  # one inconsistently oriented triangle
    _, st = convert(bs, flipped, vs, prop=True)
    assert st["sign_propagation"] == 0.0
    nan = tris.copy()
    nan[11, 4] = np.nan  # a non-finite coordinate is never "closed" (here the distance stage rejects the mesh anyway)
    with pytest.raises(bs.BsharkError):
        convert(bs, nan, vs, prop=True)


def test_fingerprint_and_exact_closedness_tests_agree(bs, bunny, monkeypatch):
    from baby_shark_b200 import synth
    meshes = [synth.config_mesh(c, s)[0] for c, s in ((3, 0.06), (4, 0.08), (5, 0.05))] + [bunny, synth.cube()]
    meshes += [m[:-3] for m in meshes[:2]]
    vss = [1 / 61, 1 / 82, 1 / 102, 0.5, 0.1, 1 / 61, 1 / 82]
    for tris, vs in zip(meshes, vss):
        res = []
        for mode in ("fingerprint", "exact"):
            monkeypatch.setenv("BSHARK_CLOSED_CHECK", mode)
            _, st = convert(bs, tris, vs, prop=True)
            res.append(st["sign_propagation"])
        monkeypatch.delenv("BSHARK_CLOSED_CHECK")
        assert res[0] == res[1], res


def test_inverted_closed_mesh(bs, oracle):
    # inward-facing orientation: winding number -1 inside, so the reference calls everything outside (wn < 0.2)
    from baby_shark_b200 import synth
    tris = synth.uv_sphere(40, 20, 0.4, (0.503, 0.504, 0.505))
    inv = tris.reshape(-1, 3, 3)[:, [0, 2, 1]].reshape(-1, 9).copy()
    v, st = convert(bs, inv, 1 / 40, prop=True)
    assert st["sign_propagation"] == 1.0
    o, _ = oracle.mesh_to_volume(inv, 1 / 40, 0, threads=8)
    compare_volumes(v.download(), o.download(), 1 / 40)
    assert v.counts()["negative"] == 0


def test_nested_and_intersecting_closed_shells(bs, oracle):
    # two closed shells in one mesh: nested (winding number 2 in the core) and partially overlapping
    from baby_shark_b200 import synth
    a = synth.uv_sphere(40, 20, 0.4, (0.503, 0.504, 0.505))
    b = synth.uv_sphere(32, 16, 0.2, (0.513, 0.494, 0.515))
    c = synth.uv_sphere(32, 16, 0.3, (0.803, 0.504, 0.505))
    for tris in (np.concatenate([a, b]), np.concatenate([a, c])):
        v, st = convert(bs, tris, 1 / 48, prop=True)
        assert st["sign_propagation"] == 1.0
        o, _ = oracle.mesh_to_volume(tris, 1 / 48, 0, threads=8)
        compare_volumes(v.download(), o.download(), 1 / 48)
