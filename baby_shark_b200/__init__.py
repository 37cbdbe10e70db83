"""baby_shark_b200 -- B200-native implicit-modelling path of baby_shark behind the C ABI in include/bshark.h.

This module is the Python harness over libbshark_cuda.so (ctypes; no torch types cross the boundary). The
classes mirror the reference's `voxel::prelude` (src/voxel/prelude.rs:1-4) and `remeshing::voxel`
(src/remeshing/voxel.rs:10-95): same names, builder-style setters, `None` where the reference returns `None`,
`ReferencePanic` where the reference panics. There is no CPU fallback: without the built library or without a
B200 the constructors raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbshark_cuda.so")

BS_OK, BS_ERR_EMPTY_MESH, BS_ERR_CUDA, BS_ERR_INVALID, BS_ERR_REFERENCE_PANICS, BS_ERR_RANGE, BS_ERR_NO_DEVICE, \
    BS_ERR_UNSUPPORTED = range(8)
_STATUS_NAMES = ["BS_OK", "BS_ERR_EMPTY_MESH", "BS_ERR_CUDA", "BS_ERR_INVALID", "BS_ERR_REFERENCE_PANICS",
                 "BS_ERR_RANGE", "BS_ERR_NO_DEVICE", "BS_ERR_UNSUPPORTED"]

BS_FLAG_COUNT_WORK, BS_FLAG_SIGN_PROPAGATION = 1, 2

# every symbol include/bshark.h declares (tests check the built library exports all of them)
EXPORTS = [
    "bs_context_create", "bs_context_destroy", "bs_last_error", "bs_context_device", "bs_context_stream",
    "bs_mesh_to_volume", "bs_mesh_to_volume_device", "bs_mesh_to_volume_sharded",
    "bs_volume_from_voxels", "bs_volume_empty", "bs_volume_sphere", "bs_volume_cuboid", "bs_volume_iwp",
    "bs_volume_clone", "bs_volume_free", "bs_volume_voxel_size",
    "bs_volume_union", "bs_volume_subtract", "bs_volume_intersect", "bs_volume_offset",
    "bs_mesh_mc", "bs_mesh_dc", "bs_mesh_mc_device", "bs_mesh_dc_device", "bs_buffer_free", "bs_voxel_remesh_into", "bs_mesh_mc_count", "bs_mesh_mc_emit_push",
    "bs_volume_download", "bs_volume_counts", "bs_context_last_stats", "bs_context_copy_out_verts", "bs_context_copy_out_verts_device", "bs_context_set_flag", "bs_kernel_launch_count",
    "bs_stl_decode", "bs_stl_decode_device", "bs_stl_encode", "bs_stl_encode_device", "bs_mesh_active_voxels", "bs_mesh_active_voxels_device",
    "bs_merge_points", "bs_merge_points_device", "bs_device_free", "bs_mesh_mc_indexed", "bs_mesh_mc_indexed_device", "bs_copy_to_host",
    "bs_ipc_alloc", "bs_ipc_open", "bs_ipc_close", "bs_ipc_free", "bs_context_push_out_verts",
]


class BsharkError(RuntimeError):
    def __init__(self, status, message=""):
        self.status = status
        name = _STATUS_NAMES[status] if 0 <= status < len(_STATUS_NAMES) else str(status)
        super().__init__("%s%s" % (name, (": " + message) if message else ""))


class ReferencePanic(BsharkError):
    """The reference panics on this input (todo!() / unwrap() / unreachable!())."""


_lib = None


def load_library(path=None):
    """Load libbshark_cuda.so (no compute). Raises if the CUDA extension has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("BSHARK_LIB") or LIB_PATH  # BSHARK_LIB: experiment builds (tools/, bench A/B runs)
    if not os.path.exists(path):
        raise ImportError("%s is missing: build it with `make -C baby_shark_b200/csrc` (there is no CPU fallback)" % path)
    L = C.CDLL(path)
    fp, vp = C.POINTER(C.c_float), C.c_void_p
    pvp = C.POINTER(C.c_void_p)
    sz = C.c_size_t
    psz = C.POINTER(sz)
    sigs = {
        "bs_context_create": (C.c_int, [C.c_int, pvp]),
        "bs_context_destroy": (None, [vp]),
        "bs_last_error": (C.c_char_p, [vp]),
        "bs_context_device": (C.c_int, [vp]),
        "bs_context_stream": (vp, [vp]),
        "bs_mesh_to_volume": (C.c_int, [vp, fp, sz, C.c_float, C.c_int64, pvp]),
        "bs_mesh_to_volume_device": (C.c_int, [vp, vp, sz, C.c_float, C.c_int64, pvp]),
        "bs_mesh_to_volume_sharded": (C.c_int, [vp, vp, sz, C.c_float, C.c_int64, C.c_int, C.c_int, pvp]),
        "bs_volume_from_voxels": (C.c_int, [vp, C.POINTER(C.c_int32), fp, sz, C.c_float, pvp]),
        "bs_volume_empty": (C.c_int, [vp, C.c_float, pvp]),
        "bs_volume_sphere": (C.c_int, [vp, C.c_float, C.c_float, fp, pvp]),
        "bs_volume_cuboid": (C.c_int, [vp, C.c_float, fp, fp, pvp]),
        "bs_volume_iwp": (C.c_int, [vp, C.c_float, fp, fp, C.c_float, pvp]),
        "bs_volume_clone": (C.c_int, [vp, pvp]),
        "bs_volume_free": (None, [vp]),
        "bs_volume_voxel_size": (C.c_float, [vp]),
        "bs_volume_union": (C.c_int, [vp, vp, pvp]),
        "bs_volume_subtract": (C.c_int, [vp, vp, pvp]),
        "bs_volume_intersect": (C.c_int, [vp, vp, pvp]),
        "bs_volume_offset": (C.c_int, [vp, C.c_float, pvp]),
        "bs_mesh_mc": (C.c_int, [vp, C.c_float, C.POINTER(fp), psz]),
        "bs_mesh_dc": (C.c_int, [vp, C.c_float, C.POINTER(fp), psz]),
        "bs_mesh_mc_device": (C.c_int, [vp, C.c_float, pvp, psz]),
        "bs_mesh_dc_device": (C.c_int, [vp, C.c_float, pvp, psz]),
        "bs_buffer_free": (None, [vp]),
        "bs_voxel_remesh_into": (C.c_int, [vp, vp, sz, C.c_float, C.c_int, C.c_int, vp, sz, psz]),
        "bs_mesh_mc_count": (C.c_int, [vp, C.c_float, psz]),
        "bs_mesh_mc_emit_push": (C.c_int, [vp, pvp, C.c_int, sz, sz]),
        "bs_volume_download": (C.c_int, [vp, C.POINTER(C.POINTER(C.c_int32)), C.POINTER(fp), C.POINTER(C.POINTER(C.c_uint64)), psz,
                                         C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.POINTER(C.c_int32)), C.POINTER(fp), psz]),
        "bs_volume_counts": (C.c_int, [vp, psz, psz, psz, psz]),
        "bs_context_last_stats": (sz, [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_double), sz]),
        "bs_context_copy_out_verts": (C.c_int, [vp, vp, sz]),
        "bs_context_copy_out_verts_device": (C.c_int, [vp, vp, sz]),
        "bs_context_set_flag": (C.c_int, [vp, C.c_int, C.c_int]),
        "bs_kernel_launch_count": (C.c_ulonglong, []),
        "bs_stl_decode": (C.c_int, [vp, vp, sz, pvp, psz]),
        "bs_stl_decode_device": (C.c_int, [vp, vp, sz, pvp, psz]),
        "bs_stl_encode": (C.c_int, [vp, vp, sz, pvp, psz]),
        "bs_stl_encode_device": (C.c_int, [vp, vp, sz, pvp, psz]),
        "bs_mesh_active_voxels": (C.c_int, [vp, pvp, psz]),
        "bs_mesh_active_voxels_device": (C.c_int, [vp, pvp, psz]),
        "bs_merge_points": (C.c_int, [vp, vp, sz, pvp, psz, pvp]),
        "bs_merge_points_device": (C.c_int, [vp, vp, sz, pvp, psz, pvp]),
        "bs_device_free": (None, [vp, vp]),
        "bs_mesh_mc_indexed": (C.c_int, [vp, C.c_float, pvp, psz, pvp, psz]),
        "bs_mesh_mc_indexed_device": (C.c_int, [vp, C.c_float, pvp, psz, pvp, psz]),
        "bs_copy_to_host": (C.c_int, [vp, vp, vp, sz]),
        "bs_ipc_alloc": (C.c_int, [vp, sz, pvp, C.c_char_p]),
        "bs_ipc_open": (C.c_int, [vp, C.c_char_p, pvp]),
        "bs_ipc_close": (C.c_int, [vp, vp]),
        "bs_ipc_free": (C.c_int, [vp, vp]),
        "bs_context_push_out_verts": (C.c_int, [vp, pvp, C.c_int, sz, sz]),
    }
    for name, (res, args) in sigs.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _lib = L
    return L


class Context:
    """One CUDA device + stream + memory pool (bs_context). `Context.default()` is the process-wide one."""
    _default = None

    def __init__(self, device=-1):
        L = load_library()
        h = C.c_void_p()
        st = L.bs_context_create(device, C.byref(h))
        if st != BS_OK:
            raise BsharkError(st, "bs_context_create(device=%d): a B200 (sm_100) device is required; there is no CPU fallback" % device)
        self._h = h

    @classmethod
    def default(cls):
        if cls._default is None:
            cls._default = Context()
        return cls._default

    def close(self):
        if getattr(self, "_h", None):
            load_library().bs_context_destroy(self._h)
            self._h = None

    def last_error(self):
        return load_library().bs_last_error(self._h).decode()

    def set_flag(self, flag, value):
        """bs_context_set_flag: BS_FLAG_COUNT_WORK = 1, BS_FLAG_SIGN_PROPAGATION = 2 (include/bshark.h)."""
        self.check(load_library().bs_context_set_flag(self._h, int(flag), int(value)))

    @property
    def device(self):
        return load_library().bs_context_device(self._h)

    @property
    def stream(self):
        return load_library().bs_context_stream(self._h)

    def last_stats(self):
        names = (C.c_char_p * 64)()
        vals = (C.c_double * 64)()
        n = load_library().bs_context_last_stats(self._h, names, vals, 64)
        return {names[i].decode(): vals[i] for i in range(n)}

    def check(self, st):
        if st == BS_OK:
            return
        if st == BS_ERR_REFERENCE_PANICS:
            raise ReferencePanic(st, self.last_error())
        raise BsharkError(st, self.last_error())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _as_triangles(mesh):
    """`Triangles::triangles()` (src/mesh/traits.rs:4-8) flattened: anything array-like of shape [n,9] / [n,3,3] /
    [3n,3] (a vertex soup, 3 consecutive vertices per triangle like PolygonSoup)."""
    a = _f32(getattr(mesh, "vertices", mesh))
    return a.reshape(-1, 9)


class Volume:
    """`voxel::volume::Volume` (src/voxel/volume/mod.rs:10-108) living on the device."""

    def __init__(self, handle, ctx):
        self._h, self._ctx = handle, ctx

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:  # at interpreter shutdown the module globals may be gone already
            _lib.bs_volume_free(self._h)
            self._h = None

    @staticmethod
    def with_voxel_size(voxel_size, ctx=None):
        ctx = ctx or Context.default()
        h = C.c_void_p()
        ctx.check(load_library().bs_volume_empty(ctx._h, voxel_size, C.byref(h)))
        return Volume(h, ctx)

    @staticmethod
    def from_fn(voxel_size, min, max, narrow_band_width, func, ctx=None):
        """volume/mod.rs:40-72. `func` maps an [m,3] f32 array of grid points to m f32 values (vectorised closure);
        it is evaluated on the host exactly as the reference does and only the kept voxels cross the boundary."""
        ctx = ctx or Context.default()
        vs = np.float32(voxel_size)
        nbw = np.float32(narrow_band_width + 1) * vs
        lo = np.floor(_f32(min) / vs).astype(np.int64)
        hi = np.ceil(_f32(max) / vs).astype(np.int64)
        xs, ys, zs = [np.arange(lo[d], hi[d] + 1, dtype=np.int64) for d in range(3)]
        ijk = np.stack(np.meshgrid(xs, ys, zs, indexing="ij"), -1).reshape(-1, 3)
        pts = ijk.astype(np.float32) * vs
        val = _f32(func(pts))
        keep = ~(np.abs(val) > nbw)
        ijk32 = np.ascontiguousarray(ijk[keep], np.int32)
        val = _f32(val[keep])
        h = C.c_void_p()
        ctx.check(load_library().bs_volume_from_voxels(ctx._h, ijk32.ctypes.data_as(C.POINTER(C.c_int32)), _fp(val), ijk32.shape[0],
                                                       voxel_size, C.byref(h)))
        return Volume(h, ctx)

    @staticmethod
    def from_voxels(ijk, values, voxel_size, ctx=None):
        """the kept voxels of a `from_fn` run, given directly: ijk [m,3] int32, values [m] f32 (later entries win)"""
        ctx = ctx or Context.default()
        ijk32, val = np.ascontiguousarray(ijk, np.int32).reshape(-1, 3), _f32(values)
        h = C.c_void_p()
        ctx.check(load_library().bs_volume_from_voxels(ctx._h, ijk32.ctypes.data_as(C.POINTER(C.c_int32)), _fp(val), ijk32.shape[0],
                                                       voxel_size, C.byref(h)))
        return Volume(h, ctx)

    def voxel_size(self):
        return load_library().bs_volume_voxel_size(self._h)

    def clone(self):
        h = C.c_void_p()
        self._ctx.check(load_library().bs_volume_clone(self._h, C.byref(h)))
        return Volume(h, self._ctx)

    def _binary(self, other, name):
        h = C.c_void_p()
        a, b = self._h, other._h
        self._h = other._h = None  # consumed (Rust move semantics)
        self._ctx.check(getattr(load_library(), "bs_volume_" + name)(a, b, C.byref(h)))
        return Volume(h, self._ctx)

    def union(self, other):
        return self._binary(other, "union")

    def subtract(self, other):
        return self._binary(other, "subtract")

    def intersect(self, other):
        return self._binary(other, "intersect")

    def offset(self, distance):
        h = C.c_void_p()
        a = self._h
        self._h = None
        self._ctx.check(load_library().bs_volume_offset(a, distance, C.byref(h)))
        return Volume(h, self._ctx)

    # parity / debug -------------------------------------------------------------------------------------
    def counts(self):
        v = [C.c_size_t() for _ in range(4)]
        self._ctx.check(load_library().bs_volume_counts(self._h, *[C.byref(x) for x in v]))
        return dict(leaves=v[0].value, active=v[1].value, negative=v[2].value, tiles=v[3].value)

    def download(self):
        L = load_library()
        ijk, tijk, tsz = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)()
        val, tval = C.POINTER(C.c_float)(), C.POINTER(C.c_float)()
        msk = C.POINTER(C.c_uint64)()
        n, nt = C.c_size_t(), C.c_size_t()
        self._ctx.check(L.bs_volume_download(self._h, C.byref(ijk), C.byref(val), C.byref(msk), C.byref(n), C.byref(tijk), C.byref(tsz),
                                             C.byref(tval), C.byref(nt)))
        n, nt = n.value, nt.value

        def take(p, count, shape, dtype):
            out = np.ctypeslib.as_array(p, shape=(max(count, 1),))[:count].astype(dtype, copy=True).reshape(shape)
            L.bs_buffer_free(C.cast(p, C.c_void_p))
            return out
        return dict(origins=take(ijk, n * 3, (-1, 3), np.int32), values=take(val, n * 512, (-1, 512), np.float32),
                    masks=take(msk, n * 8, (-1, 8), np.uint64), tile_origins=take(tijk, nt * 3, (-1, 3), np.int32),
                    tile_sizes=take(tsz, nt, (-1,), np.int32), tile_values=take(tval, nt, (-1,), np.float32))


class MeshToVolume:
    """`voxel::mesh_to_volume::MeshToVolume` (src/voxel/mesh_to_volume.rs:17-73, Default :223-236)."""

    def __init__(self, ctx=None):
        self._ctx = ctx
        self.voxel_size = 1.0
        self.band_width = 0

    def with_narrow_band_width(self, width):
        self.band_width = int(width)
        return self

    def set_narrow_band_width(self, width):
        return self.with_narrow_band_width(width)

    def with_voxel_size(self, size):
        self.voxel_size = float(size)
        return self

    def set_voxel_size(self, size):
        return self.with_voxel_size(size)

    def convert(self, mesh):
        """-> Volume, or None where the reference returns None (empty mesh, :58-60)."""
        ctx = self._ctx or Context.default()
        h = C.c_void_p()
        if isinstance(mesh, DeviceTriangles):  # decoded STL: already on the device, no host round trip
            st = load_library().bs_mesh_to_volume_device(mesh._ctx._h, mesh.ptr, mesh.n_tris, self.voxel_size, self.band_width, C.byref(h))
            ctx = mesh._ctx
        else:
            tris = _as_triangles(mesh)
            st = load_library().bs_mesh_to_volume(ctx._h, _fp(tris), tris.shape[0], self.voxel_size, self.band_width, C.byref(h))
        if st == BS_ERR_EMPTY_MESH:
            return None
        ctx.check(st)
        return Volume(h, ctx)


class VolumeBuilder:
    """`voxel::volume::builder::VolumeBuilder` (src/voxel/volume/builder.rs:5-84)."""

    def __init__(self, ctx=None):
        self._ctx = ctx
        self.voxel_size = 1.0

    def with_voxel_size(self, voxel_size):
        self.voxel_size = float(voxel_size)
        return self

    def set_voxel_size(self, voxel_size):
        self.voxel_size = float(voxel_size)

    def _make(self, fn, *args):
        ctx = self._ctx or Context.default()
        h = C.c_void_p()
        ctx.check(fn(ctx._h, self.voxel_size, *args, C.byref(h)))
        return Volume(h, ctx)

    def sphere(self, radius, origin):
        o = _f32(origin)
        return self._make(load_library().bs_volume_sphere, radius, _fp(o))

    def cuboid(self, min, max):
        a, b = _f32(min), _f32(max)
        return self._make(load_library().bs_volume_cuboid, _fp(a), _fp(b))

    def iwp(self, min, max, cell_size):
        a, b = _f32(min), _f32(max)
        return self._make(load_library().bs_volume_iwp, _fp(a), _fp(b), cell_size)


def _take_verts(p, n):
    out = np.ctypeslib.as_array(p, shape=(max(n, 1) * 3,))[: n * 3].copy().reshape(-1, 3)
    load_library().bs_buffer_free(C.cast(p, C.c_void_p))
    return out


class MarchingCubesMesher:
    """`voxel::meshing::MarchingCubesMesher` (src/voxel/meshing/marching_cubes.rs:17-63)."""

    def __init__(self):
        self.voxel_size = 1.0

    def with_voxel_size(self, size):
        self.voxel_size = float(size)
        return self

    def set_voxel_size(self, size):
        return self.with_voxel_size(size)

    def mesh(self, volume):
        """-> [n_verts,3] f32 vertex soup (3 consecutive vertices per triangle), reference emission order."""
        p, n = C.POINTER(C.c_float)(), C.c_size_t()
        volume._ctx.check(load_library().bs_mesh_mc(volume._h, self.voxel_size, C.byref(p), C.byref(n)))
        return _take_verts(p, n.value)


def mesh_indexed(volume, voxel_size):
    """Marching cubes + merge_points on the device -> IndexedVertices: the input an indexed mesh type builds from
    (src/remeshing/voxel.rs:73-83 with T = CornerTable, mesh/corner_table/builder.rs:294)."""
    pp, pi, npnt, nidx = C.c_void_p(), C.c_void_p(), C.c_size_t(), C.c_size_t()
    volume._ctx.check(load_library().bs_mesh_mc_indexed(volume._h, voxel_size, C.byref(pp), C.byref(npnt), C.byref(pi), C.byref(nidx)))
    pts = np.ctypeslib.as_array(C.cast(pp, C.POINTER(C.c_float)), shape=(max(npnt.value, 1) * 3,))[: npnt.value * 3].copy().reshape(-1, 3)
    idx = np.ctypeslib.as_array(C.cast(pi, C.POINTER(C.c_uint32)), shape=(max(nidx.value, 1),))[: nidx.value].copy()
    load_library().bs_buffer_free(pp)
    load_library().bs_buffer_free(pi)
    return IndexedVertices(pts, idx)


class DualContouringMesher:
    """`voxel::meshing::DualContouringMesher` (src/voxel/meshing/dual_contouring.rs:13-89)."""

    def __init__(self):
        self.voxel_size = 1.0

    def with_voxel_size(self, size):
        self.voxel_size = float(size)
        return self

    def mesh(self, volume):
        p, n = C.POINTER(C.c_float)(), C.c_size_t()
        volume._ctx.check(load_library().bs_mesh_dc(volume._h, self.voxel_size, C.byref(p), C.byref(n)))
        return _take_verts(p, n.value)


class MeshingMethod:
    """`remeshing::voxel::MeshingMethod` (src/remeshing/voxel.rs:10-15)."""
    FeaturePreserving = "FeaturePreserving"
    Manifold = "Manifold"


class VoxelRemesher:
    """`remeshing::voxel::VoxelRemesher` (src/remeshing/voxel.rs:45-95): convert -> MC (Manifold, default) or DC."""

    def __init__(self, ctx=None):
        self._m2v = MeshToVolume(ctx).with_narrow_band_width(0)
        self.voxel_size = 1.0
        self.meshing_method = MeshingMethod.Manifold

    def with_voxel_size(self, size):
        self._m2v.set_voxel_size(size)
        self.voxel_size = float(size)
        return self

    def with_meshing_method(self, method):
        self.meshing_method = method
        return self

    def remesh(self, mesh, slabs=0):
        """Host triangles in, host vertex soup out, through the one-call pipelined entry point (bs_voxel_remesh_into)."""
        if isinstance(mesh, DeviceTriangles):  # already on the device: convert + extract, no upload to hide
            vol = self._m2v.convert(mesh)
            if vol is None:
                return None
            if self.meshing_method == MeshingMethod.FeaturePreserving:
                return DualContouringMesher().with_voxel_size(self.voxel_size).mesh(vol)
            return MarchingCubesMesher().with_voxel_size(self.voxel_size).mesh(vol)
        tris = np.ascontiguousarray(mesh, np.float32).reshape(-1, 9)
        out = np.empty(max(1024, 4 * tris.size), np.float32)
        n = self.remesh_into(tris, out, slabs)
        if n is None:
            return None
        if n > out.size:  # the estimate was too small: the call reported the size, run again
            out = np.empty(n, np.float32)
            n = self.remesh_into(tris, out, slabs)
        return out[:n].reshape(-1, 3).copy() if n < out.size // 2 else out[:n].reshape(-1, 3)

    def remesh_into(self, tris, out, slabs=0):
        """tris: contiguous f32 array of 9 floats per triangle (host); out: contiguous f32 host array (page-locked memory
        lets the read-back overlap the kernels). Returns the number of floats of the result (larger than out.size: nothing
        usable was written, retry with that size) or None where the reference returns None."""
        ctx = self._m2v._ctx or Context.default()
        n = C.c_size_t()
        method = 1 if self.meshing_method == MeshingMethod.FeaturePreserving else 0
        st = load_library().bs_voxel_remesh_into(ctx._h, C.c_void_p(tris.ctypes.data), tris.size // 9, self.voxel_size, method, int(slabs),
                                                 C.c_void_p(out.ctypes.data), out.size, C.byref(n))
        if st == 1:
            return None
        if st == 3 and n.value > out.size:
            return n.value
        ctx.check(st)
        return n.value


class DeviceTriangles:
    """n x 9 f32 triangles resident on the context's device (library-owned): the output of `StlReader`, accepted by
    `MeshToVolume.convert`."""

    def __init__(self, ptr, n_tris, ctx):
        self.ptr, self.n_tris, self._ctx = ptr, n_tris, ctx

    def __del__(self):
        if getattr(self, "ptr", None) and _lib is not None:
            _lib.bs_device_free(self._ctx._h, self.ptr)
            self.ptr = None

    def numpy(self):
        import ctypes.util  # noqa: F401
        rt = _cudart()
        out = np.empty((self.n_tris, 9), np.float32)
        if self.n_tris:
            rt.cudaMemcpy(out.ctypes.data_as(C.c_void_p), self.ptr, out.nbytes, 2)
        return out


def _cudart():
    """libcudart as loaded by libbshark_cuda.so (only used to read DeviceTriangles back in tests)."""
    global _rt
    if _rt is None:
        import glob
        cands = glob.glob("/usr/local/cuda/lib64/libcudart.so*")
        _rt = C.CDLL(sorted(cands)[-1])
        _rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    return _rt


_rt = None


class StlReader:
    """`io::stl::StlReader` (src/io/stl.rs:14-95): binary STL -> triangles, decoded on the device."""

    def __init__(self, ctx=None):
        self._ctx = ctx

    def read_from_buffer(self, data):
        """bytes -> DeviceTriangles; raises BsharkError where the reference returns ReadError (short buffer)."""
        ctx = self._ctx or Context.default()
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        p, n = C.c_void_p(), C.c_size_t()
        ctx.check(load_library().bs_stl_decode(ctx._h, buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(p), C.byref(n)))
        return DeviceTriangles(p, n.value, ctx)

    def read_stl_from_file(self, path):
        with open(path, "rb") as f:
            return self.read_from_buffer(f.read())


class StlWriter:
    """`io::stl::StlWriter` (src/io/stl.rs:110-191) for a vertex soup: normals recomputed on the device."""

    def __init__(self, ctx=None):
        self._ctx = ctx

    def write_to_buffer(self, verts):
        """(3 n, 3) float32 vertex soup (host) -> bytes"""
        import torch
        ctx = self._ctx or Context.default()
        v = torch.from_numpy(_f32(verts).reshape(-1, 3)).cuda(ctx.device)
        p, n = C.c_void_p(), C.c_size_t()
        ctx.check(load_library().bs_stl_encode(ctx._h, C.c_void_p(v.data_ptr()), v.shape[0], C.byref(p), C.byref(n)))
        out = C.string_at(p, n.value)
        load_library().bs_buffer_free(p)
        return out


class ActiveVoxelsMesher:
    """`voxel::meshing::ActiveVoxelsMesher` (src/voxel/meshing/active_voxels.rs:4-22)."""

    def mesh(self, volume):
        """-> (m, 3) int32 vertices, three consecutive rows per triangle"""
        p, n = C.c_void_p(), C.c_size_t()
        volume._ctx.check(load_library().bs_mesh_active_voxels(volume._h, C.byref(p), C.byref(n)))
        out = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), shape=(max(n.value, 1) * 3,))[: n.value * 3].copy().reshape(-1, 3)
        load_library().bs_buffer_free(p)
        return out


class IndexedVertices:
    """`algo::merge_points::IndexedVertices` (src/algo/merge_points.rs:4-9)."""

    def __init__(self, points, indices):
        self.points, self.indices = points, indices


def merge_points(points, ctx=None):
    """`algo::merge_points::merge_points` (src/algo/merge_points.rs:12-41) on the device."""
    ctx = ctx or Context.default()
    pts = _f32(points).reshape(-1, 3)
    pu, pi, nu = C.c_void_p(), C.c_void_p(), C.c_size_t()
    ctx.check(load_library().bs_merge_points(ctx._h, pts.ctypes.data_as(C.c_void_p), pts.shape[0], C.byref(pu), C.byref(nu), C.byref(pi)))
    uniq = np.ctypeslib.as_array(C.cast(pu, C.POINTER(C.c_float)), shape=(max(nu.value, 1) * 3,))[: nu.value * 3].copy().reshape(-1, 3)
    idx = np.ctypeslib.as_array(C.cast(pi, C.POINTER(C.c_uint32)), shape=(max(pts.shape[0], 1),))[: pts.shape[0]].copy()
    load_library().bs_buffer_free(pu)
    load_library().bs_buffer_free(pi)
    return IndexedVertices(uniq, idx)


__all__ = ["Context", "Volume", "MeshToVolume", "VolumeBuilder", "MarchingCubesMesher", "DualContouringMesher",
           "VoxelRemesher", "MeshingMethod", "StlReader", "StlWriter", "DeviceTriangles", "ActiveVoxelsMesher", "IndexedVertices", "merge_points", "mesh_indexed", "BsharkError", "ReferencePanic", "load_library", "EXPORTS", "LIB_PATH"]
