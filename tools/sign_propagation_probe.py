#!/usr/bin/env python3
"""CPU probe (oracle + numpy) for the next algorithmic step of the sign stage: on a closed mesh,
are band voxels joined by "certified" links -- min(|d_p|, vs) + min(|d_q|, vs) > 1.001 |pq| -- always of the same sign, and
how many connected components (= winding-number evaluations) would remain?  Test infrastructure only."""
import os
import sys

import numpy as np
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import oracle as O  # noqa: E402
from baby_shark_b200 import synth  # noqa: E402


def mesh_is_closed(tris):
    """numpy restatement of bs_mesh_closed_impl (bs_signprop.cu): vertices identified by exact coordinates (-0 == +0),
    no triangle with a repeated vertex, every directed edge exactly once and its reverse exactly once."""
    v = np.ascontiguousarray(np.asarray(tris, np.float32).reshape(-1, 3)) + np.float32(0.0)
    u, idx = np.unique(v.view(np.uint32).reshape(-1, 3), axis=0, return_inverse=True)
    t = idx.reshape(-1, 3).astype(np.int64)
    if ((t[:, 0] == t[:, 1]) | (t[:, 1] == t[:, 2]) | (t[:, 2] == t[:, 0])).any():
        return False
    e = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]])
    k, kr = e[:, 0] * (len(u) + 1) + e[:, 1], e[:, 1] * (len(u) + 1) + e[:, 0]
    ks = np.sort(k)
    if (ks[1:] == ks[:-1]).any():
        return False
    pos = np.searchsorted(ks, kr)
    pos[pos >= ks.size] = 0
    return bool((ks[pos] == kr).all())


def probe(name, tris, vs, neighbours=6):
    d = O.mesh_to_volume(tris, vs, 0, os.cpu_count() or 1)[0].download()
    m = np.unpackbits(np.ascontiguousarray(d["masks"]).view(np.uint8).reshape(-1, 8, 8), axis=-1, bitorder="little").reshape(-1, 512).astype(bool)
    bi, off = np.nonzero(m)
    ijk = d["origins"][bi] + np.stack([off >> 6, (off >> 3) & 7, off & 7], 1)
    val = d["values"][m]
    lo = ijk.min(0)
    span = (ijk.max(0) - lo + 3).astype(np.int64)
    lin = lambda p: ((p[:, 0] - lo[0] + 1) * span[1] + (p[:, 1] - lo[1] + 1)) * span[2] + (p[:, 2] - lo[2] + 1)  # noqa: E731
    key = lin(ijk)
    order = np.argsort(key); key_s = key[order]
    cap = np.minimum(np.abs(val), np.float32(vs))
    dirs = [(1, 0, 0), (0, 1, 0), (0, 0, 1)]
    if neighbours == 18:
        dirs += [(1, 1, 0), (1, -1, 0), (1, 0, 1), (1, 0, -1), (0, 1, 1), (0, 1, -1)]
    src, dst, bad = [], [], 0
    n_links = 0
    for dx, dy, dz in dirs:
        q = ijk + np.array([dx, dy, dz])
        kq = lin(q)
        pos = np.searchsorted(key_s, kq)
        pos[pos >= key_s.size] = 0
        hit = key_s[pos] == kq
        j = order[pos]
        dist = np.float32(vs) * np.float32(np.sqrt(dx * dx + dy * dy + dz * dz))
        cert = hit & (cap + cap[j] > dist * np.float32(1.001))  # margin: a surface exactly between p and q gives d_p + d_q = |pq| up to rounding
        n_links += int(hit.sum())
        bad += int((cert & (np.signbit(val) != np.signbit(val[j]))).sum())
        src.append(np.nonzero(cert)[0]); dst.append(j[cert])
    src, dst = np.concatenate(src), np.concatenate(dst)
    g = coo_matrix((np.ones(src.size, np.int8), (src, dst)), shape=(val.size, val.size))
    ncomp, lab = connected_components(g, directed=False)
    sizes = np.bincount(lab)
    print("%-28s closed %-5s voxels %8d  links %9d certified %9d  sign conflicts on certified links %d  components %7d (largest %d, singletons %d = %.2f%% of voxels)"
          % (name, mesh_is_closed(tris), val.size, n_links, src.size, bad, ncomp, sizes.max(), int((sizes == 1).sum()), 100.0 * (sizes == 1).sum() / val.size))


if __name__ == "__main__":
    for cfg, sc in ((5, 0.06), (4, 0.08), (3, 0.06)):
        t, vs, desc = synth.config_mesh(cfg, sc)
        probe("config %d @ %g" % (cfg, sc), t, vs)
        probe("config %d @ %g (18-nbr)" % (cfg, sc), t, vs, 18)
    (ta, tb), vs, _ = synth.config_mesh(2, 0.2)
    probe("torus", ta, vs)
    bun = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "bunny_tris.npz"))["tris"]
    probe("bunny (open) @ 0.5", bun, 0.5)
