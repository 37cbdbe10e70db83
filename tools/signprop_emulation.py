#!/usr/bin/env python3
"""numpy emulation of the index arithmetic of bs_signprop.cu (k_sp_bricks / k_sp_faces / k_sp_flatten / k_sp_broadcast)
on an oracle volume: the links, the brick-face mapping and the seed-mask packing use the same formulas as the kernels.
Checks that the broadcast reproduces the oracle's signs exactly. CPU only (the kernels themselves have not run yet)."""
import os
import sys

import numpy as np
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import oracle as O  # noqa: E402
from baby_shark_b200 import synth  # noqa: E402

cfg, scale = (int(sys.argv[1]), float(sys.argv[2])) if len(sys.argv) > 2 else (5, 0.05)
tris, vs, _ = synth.config_mesh(cfg, scale)
d = O.mesh_to_volume(tris, vs, 0, os.cpu_count() or 1)[0].download()
org = d["origins"]
n = org.shape[0]
m = np.unpackbits(np.ascontiguousarray(d["masks"]).view(np.uint8).reshape(-1, 8, 8), axis=-1, bitorder="little").reshape(n, 512).astype(bool)
val = d["values"].reshape(n, 512)
vs32 = np.float32(vs)
thr = vs32 * np.float32(1.001)
cap = np.where(m, np.minimum(np.abs(val), vs32), np.float32(-1))
src, dst = [], []
t = np.arange(512)
x, y, z = t >> 6, (t >> 3) & 7, t & 7
for stride, ok in ((64, x < 7), (8, y < 7), (1, z < 7)):  # k_sp_bricks
    tt = t[ok]
    q = tt + stride
    c = m[:, tt] & (cap[:, q] >= 0) & (cap[:, tt] + cap[:, q] > thr)
    bi, k = np.nonzero(c)
    src.append(bi * 512 + tt[k]); dst.append(bi * 512 + q[k])
key = {tuple(o): i for i, o in enumerate(org // 8)}
u, v = np.repeat(np.arange(8), 8), np.tile(np.arange(8), 8)
for ax in range(3):  # k_sp_faces
    op = ((7 << 6) | (u << 3) | v) if ax == 0 else (((u << 6) | (7 << 3) | v) if ax == 1 else ((u << 6) | (v << 3) | 7))
    oq = ((u << 3) | v) if ax == 0 else (((u << 6) | v) if ax == 1 else ((u << 6) | (v << 3)))
    for b, o in enumerate(org // 8):
        nbk = list(o); nbk[ax] += 1
        nb = key.get(tuple(nbk), -1)
        if nb < 0:
            continue
        c = m[b, op] & m[nb, oq] & (cap[b, op] + cap[nb, oq] > thr)
        src.append(b * 512 + op[c]); dst.append(nb * 512 + oq[c])
        P = org[b] + np.stack([op >> 6, (op >> 3) & 7, op & 7], 1)
        Q = org[nb] + np.stack([oq >> 6, (oq >> 3) & 7, oq & 7], 1)
        e = np.zeros(3, int); e[ax] = 1
        assert (Q - P == e).all()
src, dst = np.concatenate(src), np.concatenate(dst)
N = n * 512
nc, lab = connected_components(coo_matrix((np.ones(src.size, np.int8), (src, dst)), shape=(N, N)), directed=False)
act = m.reshape(-1)
rep = np.full(nc, N, np.int64)
np.minimum.at(rep, lab[act], np.nonzero(act)[0])  # union by index: the representative is the smallest global id
par = np.where(act, rep[lab], -1)
seeds = act & (par == np.arange(N))
sgn = np.signbit(val.reshape(-1))
assert (sgn[act] == sgn[par[act]]).all()  # k_sp_broadcast reproduces the oracle
w32 = np.zeros((n, 16), np.uint32)
for tt in range(512):  # k_sp_flatten: 32-bit word t >> 5, bit t & 31
    w32[:, tt >> 5] |= seeds.reshape(n, 512)[:, tt].astype(np.uint32) << np.uint32(tt & 31)
chk = np.unpackbits(w32.view(np.uint64).view(np.uint8).reshape(n, 8, 8), axis=-1, bitorder="little").reshape(n, 512).astype(bool)
assert (chk == seeds.reshape(n, 512)).all()
print("config %d @ %g: %d bricks, %d active voxels, %d representatives (%.2f %%), %d bricks without one: broadcast reproduces the oracle signs"
      % (cfg, scale, n, act.sum(), seeds.sum(), 100.0 * seeds.sum() / act.sum(), int((seeds.reshape(n, 512).sum(1) == 0).sum())))
