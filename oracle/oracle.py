"""ORACLE -- TEST INFRASTRUCTURE ONLY. ctypes wrapper over oracle/libbs_oracle.so (CPU restatement of the
reference's implicit-modelling path, see bs_oracle.cpp). Imported only by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs; never by baby_shark_b200/.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libbs_oracle.so")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class ConvertStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_tris", "n_sub", "n_eval", "n_active", "n_leaves", "n_negative",
                                          "wn_visit", "wn_far", "wn_exact", "tree_nodes")] + \
               [(n, C.c_double) for n in ("t_subdivide", "t_tree", "t_udf", "t_sign")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class SweepStats(C.Structure):
    _fields_ = [("leaves_processed", C.c_uint64 * 8), ("n_leaves_final", C.c_uint64)]


class McStats(C.Structure):
    _fields_ = [("case_hist", C.c_uint64 * 15), ("n_cubes", C.c_uint64), ("n_degenerate", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp = C.POINTER(C.c_float)
        L.bso_mesh_to_volume.restype = C.c_void_p
        L.bso_mesh_to_volume.argtypes = [fp, C.c_size_t, C.c_float, C.c_int64, C.c_int, C.POINTER(ConvertStats), C.c_int]
        L.bso_volume_sphere.restype = C.c_void_p
        L.bso_volume_sphere.argtypes = [C.c_float] * 5
        L.bso_volume_cuboid.restype = C.c_void_p
        L.bso_volume_cuboid.argtypes = [C.c_float, fp, fp]
        L.bso_volume_iwp.restype = C.c_void_p
        L.bso_volume_iwp.argtypes = [C.c_float, fp, fp, C.c_float]
        L.bso_volume_empty.restype = C.c_void_p
        L.bso_volume_empty.argtypes = [C.c_float]
        L.bso_volume_from_voxels.restype = C.c_void_p
        L.bso_volume_from_voxels.argtypes = [C.POINTER(C.c_int32), fp, C.c_size_t, C.c_float]
        L.bso_volume_clone.restype = C.c_void_p
        L.bso_volume_clone.argtypes = [C.c_void_p]
        L.bso_volume_free.argtypes = [C.c_void_p]
        L.bso_volume_voxel_size.restype = C.c_float
        L.bso_volume_voxel_size.argtypes = [C.c_void_p]
        for n in ("union", "subtract", "intersect"):
            getattr(L, "bso_volume_" + n).argtypes = [C.c_void_p, C.c_void_p]
            getattr(L, "bso_volume_" + n).restype = None
        L.bso_volume_flood_fill.argtypes = [C.c_void_p]
        L.bso_volume_sign_at.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64]
        L.bso_volume_offset.argtypes = [C.c_void_p, C.c_float, C.POINTER(SweepStats)]
        L.bso_volume_counts.argtypes = [C.c_void_p] + [C.POINTER(C.c_size_t)] * 4
        L.bso_volume_download.argtypes = [C.c_void_p, C.POINTER(C.c_int32), fp, C.POINTER(C.c_uint64),
                                          C.POINTER(C.c_int32), C.POINTER(C.c_int32), fp]
        L.bso_mesh_mc.argtypes = [C.c_void_p, C.c_float, C.POINTER(fp), C.POINTER(C.c_size_t), C.POINTER(McStats)]
        L.bso_mesh_dc.argtypes = [C.c_void_p, C.c_float, C.POINTER(fp), C.POINTER(C.c_size_t)]
        L.bso_buffer_free.argtypes = [fp]
        L.bso_subdivide.restype = C.c_size_t
        L.bso_subdivide.argtypes = [fp, C.c_size_t, C.c_float, C.POINTER(fp)]
        L.bso_point_triangle_distance.argtypes = [fp, fp, C.c_size_t, fp]
        L.bso_winding_numbers.argtypes = [fp, C.c_size_t, fp, C.c_size_t, C.c_float, fp, C.POINTER(C.c_uint64)]
        L.bso_compute_distance.restype = C.c_float
        L.bso_compute_distance.argtypes = [C.c_float] * 4
        L.bso_selftest.restype = C.c_int
        u8p = C.POINTER(C.c_uint8)
        L.bso_stl_decode.argtypes = [u8p, C.c_size_t, C.POINTER(fp), C.POINTER(C.c_size_t)]
        L.bso_stl_encode.argtypes = [fp, C.c_size_t, u8p]
        L.bso_stl_encode.restype = None
        L.bso_active_voxels.restype = C.c_size_t
        L.bso_active_voxels.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_int32))]
        L.bso_merge_points.restype = C.c_size_t
        L.bso_merge_points.argtypes = [fp, C.c_size_t, fp, C.POINTER(C.c_uint32)]
        L.bso_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Volume:
    """Owns one oracle volume (reference `voxel::volume::Volume`, volume/mod.rs:10-14)."""

    def __init__(self, handle):
        assert handle
        self._h = handle

    def __del__(self):
        if getattr(self, "_h", None):
            lib().bso_volume_free(self._h)
            self._h = None

    @property
    def voxel_size(self):
        return lib().bso_volume_voxel_size(self._h)

    def clone(self):
        return Volume(lib().bso_volume_clone(self._h))

    def _consume(self, other, name):
        getattr(lib(), "bso_volume_" + name)(self._h, other._h)
        other._h = None
        return self

    def union(self, other):
        return self._consume(other, "union")

    def subtract(self, other):
        return self._consume(other, "subtract")

    def intersect(self, other):
        return self._consume(other, "intersect")

    def offset(self, distance):
        st = SweepStats()
        rc = lib().bso_volume_offset(self._h, distance, C.byref(st))
        if rc:
            raise RuntimeError("offset: the reference panics on this input")
        self.sweep_stats = st
        return self

    def counts(self):
        v = [C.c_size_t() for _ in range(4)]
        lib().bso_volume_counts(self._h, *[C.byref(x) for x in v])
        return dict(leaves=v[0].value, active=v[1].value, tiles=v[2].value, negative=v[3].value)

    def download(self):
        """-> dict(origins[n,3] i32, values[n,512] f32, masks[n,8] u64 (bit o&63 of word o>>6, o = x<<6|y<<3|z),
        tile_origins[t,3], tile_sizes[t], tile_values[t]); leaves and tiles in the reference's visit order."""
        c = self.counts()
        n, t = c["leaves"], c["tiles"]
        origins = np.zeros((max(n, 1), 3), np.int32)
        values = np.zeros((max(n, 1), 512), np.float32)
        masks = np.zeros((max(n, 1), 8), np.uint64)
        to = np.zeros((max(t, 1), 3), np.int32)
        ts = np.zeros(max(t, 1), np.int32)
        tv = np.zeros(max(t, 1), np.float32)
        lib().bso_volume_download(self._h, origins.ctypes.data_as(C.POINTER(C.c_int32)), _fp(values),
                                  masks.ctypes.data_as(C.POINTER(C.c_uint64)), to.ctypes.data_as(C.POINTER(C.c_int32)),
                                  ts.ctypes.data_as(C.POINTER(C.c_int32)), _fp(tv))
        return dict(origins=origins[:n], values=values[:n], masks=masks[:n], tile_origins=to[:t], tile_sizes=ts[:t],
                    tile_values=tv[:t])

    def sign_at(self, x, y, z):
        return lib().bso_volume_sign_at(self._h, x, y, z)


def mesh_to_volume(tris, voxel_size, band=0, threads=1, count_work=False):
    """reference `MeshToVolume::convert` (mesh_to_volume.rs:52-73); returns (Volume | None, stats dict).
    count_work: also fill the wn_* traversal counters (never set in a timed run)."""
    tris = _f32(tris).reshape(-1, 9)
    st = ConvertStats()
    h = lib().bso_mesh_to_volume(_fp(tris), tris.shape[0], voxel_size, band, threads, C.byref(st), int(bool(count_work)))
    return (Volume(h) if h else None), st.as_dict()


def sphere(voxel_size, radius, origin):
    return Volume(lib().bso_volume_sphere(voxel_size, radius, *[float(x) for x in origin]))


def cuboid(voxel_size, mn, mx):
    mn, mx = _f32(mn), _f32(mx)
    return Volume(lib().bso_volume_cuboid(voxel_size, _fp(mn), _fp(mx)))


def iwp(voxel_size, mn, mx, cell_size):
    mn, mx = _f32(mn), _f32(mx)
    return Volume(lib().bso_volume_iwp(voxel_size, _fp(mn), _fp(mx), cell_size))


def empty(voxel_size):
    return Volume(lib().bso_volume_empty(voxel_size))


def from_voxels(ijk, values, voxel_size):
    ijk = np.ascontiguousarray(ijk, np.int32).reshape(-1, 3)
    values = _f32(values)
    return Volume(lib().bso_volume_from_voxels(ijk.ctypes.data_as(C.POINTER(C.c_int32)), _fp(values), ijk.shape[0], voxel_size))


def _take(ptr, n):
    out = np.ctypeslib.as_array(ptr, shape=(max(n, 1) * 3,))[: n * 3].copy().reshape(-1, 3)
    lib().bso_buffer_free(ptr)
    return out


def marching_cubes(vol, voxel_size=None, with_stats=False):
    """reference `MarchingCubesMesher::mesh` (marching_cubes.rs:43-63) -> [n_verts,3] f32 in emission order."""
    p = C.POINTER(C.c_float)()
    n = C.c_size_t()
    st = McStats()
    lib().bso_mesh_mc(vol._h, vol.voxel_size if voxel_size is None else voxel_size, C.byref(p), C.byref(n), C.byref(st))
    v = _take(p, n.value)
    return (v, st) if with_stats else v


def dual_contouring(vol, voxel_size=None):
    """reference `DualContouringMesher::mesh`; None where the reference returns None / panics."""
    p = C.POINTER(C.c_float)()
    n = C.c_size_t()
    rc = lib().bso_mesh_dc(vol._h, vol.voxel_size if voxel_size is None else voxel_size, C.byref(p), C.byref(n))
    if rc:
        return None
    return _take(p, n.value)


def subdivide(tris, voxel_size):
    tris = _f32(tris).reshape(-1, 9)
    p = C.POINTER(C.c_float)()
    n = lib().bso_subdivide(_fp(tris), tris.shape[0], voxel_size, C.byref(p))
    out = np.ctypeslib.as_array(p, shape=(max(n, 1) * 9,))[: n * 9].copy().reshape(-1, 9)
    lib().bso_buffer_free(p)
    return out


def point_triangle_distance(tri, pts):
    tri, pts = _f32(tri).reshape(9), _f32(pts).reshape(-1, 3)
    out = np.zeros(pts.shape[0], np.float32)
    lib().bso_point_triangle_distance(_fp(tri), _fp(pts), pts.shape[0], _fp(out))
    return out


def winding_numbers(tris, pts, beta=2.0):
    tris, pts = _f32(tris).reshape(-1, 9), _f32(pts).reshape(-1, 3)
    out = np.zeros(pts.shape[0], np.float32)
    cnt = (C.c_uint64 * 4)()
    lib().bso_winding_numbers(_fp(tris), tris.shape[0], _fp(pts), pts.shape[0], beta, _fp(out), cnt)
    return out, dict(visit=cnt[0], far=cnt[1], exact=cnt[2], nodes=cnt[3])


def stl_decode(data):
    """io/stl.rs StlReader: bytes -> (n, 9) float32 triangles, or None on a short buffer (the reference returns ReadError)."""
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    out, n = C.POINTER(C.c_float)(), C.c_size_t()
    if lib().bso_stl_decode(buf.ctypes.data_as(C.POINTER(C.c_uint8)), buf.size, C.byref(out), C.byref(n)):
        return None
    a = np.ctypeslib.as_array(out, shape=(max(1, n.value * 9),))[: n.value * 9].copy().reshape(-1, 9)
    lib().bso_free(out)
    return a


def stl_encode(verts):
    """io/stl.rs StlWriter on a triangle soup ((3 n, 3) float32 vertices) -> bytes."""
    v = _f32(verts).reshape(-1, 9)
    out = np.zeros(84 + 50 * v.shape[0], np.uint8)
    lib().bso_stl_encode(_fp(v), v.shape[0], out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out.tobytes()


def active_voxels(vol):
    """ActiveVoxelsMesher::mesh -> (m, 3) int32 vertices, three consecutive rows per triangle."""
    out = C.POINTER(C.c_int32)()
    n = lib().bso_active_voxels(vol._h, C.byref(out))
    a = np.ctypeslib.as_array(out, shape=(max(1, n * 3),))[: n * 3].copy().reshape(-1, 3)
    lib().bso_free(out)
    return a


def merge_points(points):
    """algo::merge_points -> (unique (k, 3) float32 in first-occurrence order, indices (n,) uint32)."""
    p = _f32(points).reshape(-1, 3)
    uniq = np.zeros((max(1, p.shape[0]), 3), np.float32)
    idx = np.zeros(max(1, p.shape[0]), np.uint32)
    k = lib().bso_merge_points(_fp(p), p.shape[0], _fp(uniq), idx.ctypes.data_as(C.POINTER(C.c_uint32)))
    return uniq[:k].copy(), idx[: p.shape[0]].copy()


def selftest():
    return lib().bso_selftest()
