"""Open input meshes (-m gpu): the sign of a voxel is `winding number < 0.2` (mesh_to_volume.rs:264-271) with the winding
number APPROXIMATED over the reference's SAH tree (aabb_tree.rs:654-691, beta = 2). The device walks an LBVH (north_star), so
its far-field approximations are taken over other node sets: topology and |SDF| stay bit-identical, the sign can differ
where the winding number passes 0.2, which on an open mesh happens along a surface that leaves the hole. These tests
measure how often, and pin every disagreement to that surface: the EXACT winding number there is within 0.1 of the
threshold (the reference's own approximation error at beta = 2 is up to 0.06)."""
import numpy as np
import pytest

from baby_shark_b200 import synth
from util import active_mask_bits

pytestmark = pytest.mark.gpu


def open_cases(bunny):
    s = synth.uv_sphere(96, 48, 0.4, (0.503, 0.504, 0.505))
    cz = s.reshape(-1, 3, 3)[:, :, 2].mean(1)
    yield "sphere without its top cap", s[cz < 0.78], 1.0 / 64
    yield "half sphere", s[cz < 0.505], 1.0 / 64
    by = bunny.reshape(-1, 3, 3)[:, :, 1].mean(1)
    lo, hi = by.min(), by.max()
    yield "bunny with the top 15 % removed", bunny[by < lo + 0.85 * (hi - lo)], 0.5


def test_open_mesh_sign_disagreements_sit_at_the_threshold(bs, oracle, bunny):
    report = []
    for name, tris, vs in open_cases(bunny):
        tris = np.ascontiguousarray(tris, np.float32)
        g = bs.MeshToVolume().with_voxel_size(vs).convert(tris)
        assert bs.Context.default().last_stats()["sign_propagation"] == 0.0  # open: the per-voxel path
        g = g.download()
        o = oracle.mesh_to_volume(tris, vs, 0, 8)[0].download()
        assert np.array_equal(g["origins"], o["origins"])
        m = active_mask_bits(o["masks"])
        assert np.array_equal(active_mask_bits(g["masks"]), m)
        gv, ov = g["values"][m], o["values"][m]
        assert np.array_equal(np.abs(gv).view(np.uint32), np.abs(ov).view(np.uint32))
        diff = np.signbit(gv) != np.signbit(ov)
        n_bad, frac = int(diff.sum()), float(diff.mean())
        worst = 0.0
        if n_bad:
            idx = np.argwhere(m)[diff]
            pts = (o["origins"][idx[:, 0]] + np.stack([idx[:, 1] >> 6, (idx[:, 1] >> 3) & 7, idx[:, 1] & 7], 1)).astype(np.float32) * np.float32(vs)
            wn_exact, _ = oracle.winding_numbers(tris, pts[:2000], beta=-1.0)
            worst = float(np.abs(wn_exact - 0.2).max())
            assert worst < 0.1, (name, worst)
        report.append("%s: %d active voxels, %d sign disagreements (%.4f %%), worst |wn_exact - 0.2| = %.3f" % (name, gv.size, n_bad, 100 * frac, worst))
        assert frac < 5e-3, report[-1]
    print("\n".join(report))
