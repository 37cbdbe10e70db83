"""Deterministic synthetic input meshes for the benchmark configs (SURVEY.md section 8d / BASELINE.md).

All generators return an [n,9] float32 array (p1,p2,p3 per triangle), outward orientation, closed surfaces.
The same arrays can be written as binary STL (`write_binary_stl`) so a Rust run of the reference can consume the
same bytes.
"""
import numpy as np


def _grid_quads_to_tris(P, wrap_u=True, wrap_v=False, flip=False):
    """P: [nu, nv, 3] vertex grid. Two triangles per quad, counter-clockwise seen from outside for a sphere/torus
    parametrised with u = longitude (wrapping) and v = the other angle."""
    nu, nv = P.shape[0], P.shape[1]
    iu = np.arange(nu if wrap_u else nu - 1)
    iv = np.arange(nv if wrap_v else nv - 1)
    U, V = np.meshgrid(iu, iv, indexing="ij")
    U1, V1 = (U + 1) % nu, (V + 1) % nv
    a, b, c, d = P[U, V], P[U1, V], P[U1, V1], P[U, V1]
    if flip:
        b, d = d, b
    t1 = np.concatenate([a, b, c], -1).reshape(-1, 9)
    t2 = np.concatenate([a, c, d], -1).reshape(-1, 9)
    return np.concatenate([t1, t2], 0)


def uv_sphere(n_lon, n_lat, radius, center, displace=None):
    """UV sphere: n_lon longitudes x n_lat latitude bands; the two pole bands are single triangles
    (2*n_lon*(n_lat-1) triangles). `displace(unit_dirs[m,3]) -> radial offsets[m]` keeps it star-shaped."""
    center = np.asarray(center, np.float64)
    lon = np.arange(n_lon) * (2.0 * np.pi / n_lon)
    lat = np.arange(1, n_lat) * (np.pi / n_lat)  # interior rings
    LON, LAT = np.meshgrid(lon, lat, indexing="ij")
    D = np.stack([np.sin(LAT) * np.cos(LON), np.sin(LAT) * np.sin(LON), np.cos(LAT)], -1)
    poles = np.array([[0.0, 0.0, 1.0], [0.0, 0.0, -1.0]])

    def place(dirs):
        r = radius + (displace(dirs.reshape(-1, 3)).reshape(dirs.shape[:-1]) if displace is not None else 0.0)
        return center + dirs * np.asarray(r)[..., None]
    P = place(D)
    north, south = place(poles[0:1])[0], place(poles[1:2])[0]
    body = _grid_quads_to_tris(P, wrap_u=True, wrap_v=False, flip=True)
    nxt = (np.arange(n_lon) + 1) % n_lon
    cap_n = np.concatenate([np.broadcast_to(north, (n_lon, 3)), P[:, 0], P[nxt, 0]], -1)
    cap_s = np.concatenate([np.broadcast_to(south, (n_lon, 3)), P[nxt, -1], P[:, -1]], -1)
    return np.ascontiguousarray(np.concatenate([cap_n, body, cap_s], 0), np.float32)


def torus(n_u, n_v, R, r, center, rotate_x_90=False):
    """UV torus, n_u x n_v quads -> 2*n_u*n_v triangles; axis z (or y after the 90 degree rotation about x)."""
    u = np.arange(n_u) * (2.0 * np.pi / n_u)
    v = np.arange(n_v) * (2.0 * np.pi / n_v)
    U, V = np.meshgrid(u, v, indexing="ij")
    P = np.stack([(R + r * np.cos(V)) * np.cos(U), (R + r * np.cos(V)) * np.sin(U), r * np.sin(V)], -1)
    if rotate_x_90:
        P = np.stack([P[..., 0], -P[..., 2], P[..., 1]], -1)
    P = P + np.asarray(center, np.float64)
    return np.ascontiguousarray(_grid_quads_to_tris(P, wrap_u=True, wrap_v=True), np.float32)


def lattice_noise(seed=0, octaves=4):
    """Sum of `octaves` sin-product lattice-noise octaves with seed-`seed` phases, in [-1, 1]."""
    rng = np.random.default_rng(seed)
    phases = rng.uniform(0.0, 2.0 * np.pi, size=(octaves, 3))

    def f(dirs):
        out = np.zeros(dirs.shape[0])
        amp, freq, tot = 1.0, 4.0, 0.0
        for o in range(octaves):
            out += amp * np.sin(freq * dirs[:, 0] + phases[o, 0]) * np.sin(freq * dirs[:, 1] + phases[o, 1]) * np.sin(freq * dirs[:, 2] + phases[o, 2])
            tot += amp
            amp *= 0.5
            freq *= 2.0
        return out / tot
    return f


def cube(center=(0.0, 0.0, 0.0), sx=1.0, sy=1.0, sz=1.0):
    """Axis-aligned box, 12 triangles (the shape `mesh::builder::cube` builds for remeshing/voxel.rs:105-112)."""
    c = np.asarray(center, np.float64)
    h = np.array([sx, sy, sz]) * 0.5
    v = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], np.float64) * h + c
    f = [(0, 2, 1), (0, 3, 2), (4, 5, 6), (4, 6, 7), (0, 1, 5), (0, 5, 4), (1, 2, 6), (1, 6, 5), (2, 3, 7), (2, 7, 6), (3, 0, 4), (3, 4, 7)]
    return np.ascontiguousarray(np.array([[*v[a], *v[b], *v[c_]] for a, b, c_ in f]), np.float32)


OFF_GRID = np.array([0.3, 0.4, 0.5])  # avoids grid alignment


def config_mesh(cfg, scale=1.0):
    """Benchmark config -> (tris, voxel_size, description). `scale` < 1 shrinks the resolution (and triangle
    count quadratically) for parity tests; scale = 1 is the BASELINE.json size."""
    if cfg == 1:  # examples/voxel_remeshing.rs on assets/bunny.stl (13 000 triangles, closed), voxel_size 0.01
        import os
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "bunny_tris.npz")
        vs = 0.01 / scale
        return np.ascontiguousarray(np.load(path)["tris"], np.float32), vs, "assets/bunny.stl, 13000 triangles, voxel %g (~%d^3 grid)" % (vs, int(52 / vs))
    if cfg == 2:
        res = max(32, int(round(512 * scale)))
        vs = 1.0 / res
        nu, nv = max(16, int(1024 * scale)), max(8, int(512 * scale))
        c = np.full(3, 0.5) + OFF_GRID * vs
        a = torus(nu, nv, 0.30, 0.10, c)
        b = torus(nu, nv, 0.30, 0.10, c + np.array([0.0, 0.30, 0.0]), rotate_x_90=True)
        return (a, b), vs, "two UV tori %dx%d quads, R=0.30 r=0.10, grid %d^3" % (nu, nv, res)
    if cfg == 3:
        res = max(32, int(round(1024 * scale)))
        vs = 1.0 / res
        nlon, nlat = max(16, int(1000 * scale)), max(8, int(500 * scale))
        return uv_sphere(nlon, nlat, 0.44, np.full(3, 0.5) + OFF_GRID * vs), vs, "UV sphere %dx%d, R=0.44, grid %d^3" % (nlon, nlat, res)
    if cfg == 4:
        res = max(32, int(round(1024 * scale)))
        vs = 1.0 / res
        nlon, nlat = max(16, int(2000 * scale)), max(8, int(500 * scale))
        noise = lattice_noise(0)
        return uv_sphere(nlon, nlat, 0.40, np.full(3, 0.5) + OFF_GRID * vs, lambda d: 0.02 * noise(d)), vs, \
            "noise-displaced UV sphere %dx%d, R=0.40, grid %d^3" % (nlon, nlat, res)
    if cfg == 5:
        res = max(32, int(round(2048 * scale)))
        vs = 1.0 / res
        nlon, nlat = max(16, int(3162 * scale)), max(8, int(1581 * scale))
        noise = lattice_noise(0)
        return uv_sphere(nlon, nlat, 0.44, np.full(3, 0.5) + OFF_GRID * vs, lambda d: 0.02 * noise(d)), vs, \
            "noise-displaced UV sphere %dx%d, R=0.44, grid %d^3" % (nlon, nlat, res)
    raise ValueError(cfg)


def write_binary_stl(path, tris):
    tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
    rec = np.zeros((tris.shape[0], 50), np.uint8)
    rec[:, 12:48] = tris.view(np.uint8).reshape(-1, 36)
    with open(path, "wb") as f:
        f.write(b"\0" * 80)
        f.write(np.uint32(tris.shape[0]).tobytes())
        f.write(rec.tobytes())


def read_binary_stl(path):
    """io/stl.rs:65-95: normals ignored, vertices in file order."""
    b = open(path, "rb").read()
    n = int(np.frombuffer(b[80:84], np.uint32)[0])
    rec = np.frombuffer(b[84:84 + 50 * n], dtype=np.uint8).reshape(n, 50)
    return rec[:, 12:48].copy().view("<f4").reshape(n, 9)
