//! `voxel::prelude` of baby_shark (src/voxel/prelude.rs:1-4) with the same public signatures, backed by the B200
//! library through `ffi`. Swap `use baby_shark::voxel::prelude::*` for `use baby_shark_voxel_b200::prelude::*`,
//! or re-export this crate as `baby_shark::voxel` behind a cargo feature.
//! Source only: this file has never been compiled (no Rust toolchain in the build container).
mod ffi;
use nalgebra::Vector3;
use std::ptr;
use std::sync::OnceLock;

pub type Vec3f = Vector3<f32>;

struct Ctx(*mut ffi::bs_context);
// The library serialises every call on a context with its own lock (bs_common.cuh BS_ENTER), so the handle may be shared.
// A `*_device` extraction leaves its result in a per-context buffer until the next extraction: EXTRACT keeps
// "extract, then copy out" one critical section across threads.
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}
static EXTRACT: std::sync::Mutex<()> = std::sync::Mutex::new(());
fn ctx() -> *mut ffi::bs_context {
    static CTX: OnceLock<Ctx> = OnceLock::new();
    CTX.get_or_init(|| {
        let mut h = ptr::null_mut();
        let st = unsafe { ffi::bs_context_create(-1, &mut h) };
        assert!(st == ffi::BS_OK, "bshark: no B200 available (status {st}); there is no CPU fallback");
        Ctx(h)
    }).0
}
fn check(st: ffi::bs_status) {
    if st != ffi::BS_OK {
        let msg = unsafe { std::ffi::CStr::from_ptr(ffi::bs_last_error(ctx())) }.to_string_lossy().into_owned();
        // BS_ERR_REFERENCE_PANICS: the reference panics here too (todo!() / unwrap() / unreachable!())
        panic!("bshark status {st}: {msg}");
    }
}

/// voxel::volume::Volume (src/voxel/volume/mod.rs:10-123); lives on the device.
#[derive(Debug)]
pub struct Volume { h: *mut ffi::bs_volume }
unsafe impl Send for Volume {}
impl Drop for Volume { fn drop(&mut self) { unsafe { ffi::bs_volume_free(self.h) } } }
impl Clone for Volume {
    fn clone(&self) -> Self { let mut h = ptr::null_mut(); check(unsafe { ffi::bs_volume_clone(self.h, &mut h) }); Self { h } }
}
impl Volume {
    pub fn with_voxel_size(voxel_size: f32) -> Self { let mut h = ptr::null_mut(); check(unsafe { ffi::bs_volume_empty(ctx(), voxel_size, &mut h) }); Self { h } }
    pub fn voxel_size(&self) -> f32 { unsafe { ffi::bs_volume_voxel_size(self.h) } }
    /// volume/mod.rs:40-72 -- the closure runs on the host exactly as in the reference; kept voxels cross the FFI.
    pub fn from_fn<TFn: Fn(&Vec3f) -> f32>(voxel_size: f32, min: Vec3f, max: Vec3f, narrow_band_width: usize, func: TFn) -> Self {
        let nbw = (narrow_band_width + 1) as f32 * voxel_size;
        let lo = (min / voxel_size).map(|x| x.floor() as isize);
        let hi = (max / voxel_size).map(|x| x.ceil() as isize);
        let (mut ijk, mut val) = (Vec::<i32>::new(), Vec::<f32>::new());
        for x in lo.x..=hi.x { for y in lo.y..=hi.y { for z in lo.z..=hi.z {
            let v = func(&(Vector3::new(x as f32, y as f32, z as f32) * voxel_size));
            if v.abs() > nbw { continue; }
            ijk.extend_from_slice(&[x as i32, y as i32, z as i32]); val.push(v);
        } } }
        let mut h = ptr::null_mut();
        check(unsafe { ffi::bs_volume_from_voxels(ctx(), ijk.as_ptr(), val.as_ptr(), val.len(), voxel_size, &mut h) });
        Self { h }
    }
    fn binary(self, other: Self, f: unsafe extern "C" fn(*mut ffi::bs_volume, *mut ffi::bs_volume, *mut *mut ffi::bs_volume) -> ffi::bs_status) -> Self {
        let (a, b) = (self.h, other.h);
        std::mem::forget(self); std::mem::forget(other);  // the library consumes both handles
        let mut h = ptr::null_mut();
        check(unsafe { f(a, b, &mut h) });
        Self { h }
    }
    pub fn union(self, other: Self) -> Self { self.binary(other, ffi::bs_volume_union) }
    pub fn intersect(self, other: Self) -> Self { self.binary(other, ffi::bs_volume_intersect) }
    pub fn subtract(self, other: Self) -> Self { self.binary(other, ffi::bs_volume_subtract) }
    pub fn offset(self, distance: f32) -> Self {
        let a = self.h; std::mem::forget(self);
        let mut h = ptr::null_mut();
        check(unsafe { ffi::bs_volume_offset(a, distance, &mut h) });
        Self { h }
    }
}

/// voxel::mesh_to_volume::MeshToVolume (src/voxel/mesh_to_volume.rs:17-73)
pub struct MeshToVolume { band_width: isize, voxel_size: f32 }
impl Default for MeshToVolume { fn default() -> Self { Self { band_width: 0, voxel_size: 1.0 } } }
impl MeshToVolume {
    pub fn with_narrow_band_width(mut self, width: isize) -> Self { self.band_width = width; self }
    pub fn set_narrow_band_width(&mut self, width: isize) -> &mut Self { self.band_width = width; self }
    pub fn with_voxel_size(mut self, size: f32) -> Self { self.voxel_size = size; self }
    pub fn set_voxel_size(&mut self, size: f32) -> *mut Self { self.voxel_size = size; self }
    /// `mesh.triangles()` is flattened to n x 9 f32 (mesh/traits.rs:4-8). In-tree this takes `&T where T: Triangles<Scalar = f32>`.
    pub fn convert<I: Iterator<Item = [Vec3f; 3]>>(&mut self, triangles: I) -> Option<Volume> {
        let flat: Vec<f32> = triangles.flat_map(|t| t.into_iter().flat_map(|p| [p.x, p.y, p.z])).collect();
        let mut h = ptr::null_mut();
        let st = unsafe { ffi::bs_mesh_to_volume(ctx(), flat.as_ptr(), flat.len() / 9, self.voxel_size, self.band_width as i64, &mut h) };
        if st == ffi::BS_ERR_EMPTY_MESH { return None; }
        check(st);
        Some(Volume { h })
    }
}

/// voxel::volume::builder::VolumeBuilder (src/voxel/volume/builder.rs:5-84)
pub struct VolumeBuilder { voxel_size: f32 }
impl Default for VolumeBuilder { fn default() -> Self { Self { voxel_size: 1.0 } } }
impl VolumeBuilder {
    pub fn with_voxel_size(mut self, voxel_size: f32) -> Self { self.voxel_size = voxel_size; self }
    pub fn set_voxel_size(&mut self, voxel_size: f32) { self.voxel_size = voxel_size; }
    pub fn sphere(&self, radius: f32, origin: Vec3f) -> Volume { let mut h = ptr::null_mut(); check(unsafe { ffi::bs_volume_sphere(ctx(), self.voxel_size, radius, origin.as_ptr(), &mut h) }); Volume { h } }
    pub fn cuboid(&self, min: Vec3f, max: Vec3f) -> Volume { let mut h = ptr::null_mut(); check(unsafe { ffi::bs_volume_cuboid(ctx(), self.voxel_size, min.as_ptr(), max.as_ptr(), &mut h) }); Volume { h } }
    pub fn iwp(&self, min: Vec3f, max: Vec3f, cell_size: f32) -> Volume { let mut h = ptr::null_mut(); check(unsafe { ffi::bs_volume_iwp(ctx(), self.voxel_size, min.as_ptr(), max.as_ptr(), cell_size, &mut h) }); Volume { h } }
}

fn take_vertices(n_verts: usize) -> Vec<Vec3f> {
    let mut out: Vec<Vec3f> = Vec::with_capacity(n_verts);
    // Vector3<f32> is repr(C) [f32; 3]: the device result is copied straight into the Vec the caller receives
    check(unsafe { ffi::bs_context_copy_out_verts(ctx(), out.as_mut_ptr() as *mut f32, n_verts * 3) });
    unsafe { out.set_len(n_verts) };
    out
}

/// voxel::meshing::MarchingCubesMesher (src/voxel/meshing/marching_cubes.rs:17-63)
pub struct MarchingCubesMesher { voxel_size: f32 }
impl Default for MarchingCubesMesher { fn default() -> Self { Self { voxel_size: 1.0 } } }
impl MarchingCubesMesher {
    pub fn with_voxel_size(mut self, size: f32) -> Self { self.voxel_size = size; self }
    pub fn set_voxel_size(&mut self, size: f32) -> &mut Self { self.voxel_size = size; self }
    pub fn mesh(&mut self, sdf: &Volume) -> Vec<Vec3f> {
        let _pair = EXTRACT.lock().unwrap_or_else(|e| e.into_inner());
        let (mut d, mut n) = (ptr::null(), 0usize);
        check(unsafe { ffi::bs_mesh_mc_device(sdf.h, self.voxel_size, &mut d, &mut n) });
        take_vertices(n)
    }
}

/// voxel::meshing::DualContouringMesher (src/voxel/meshing/dual_contouring.rs:13-89)
pub struct DualContouringMesher { voxel_size: f32 }
impl Default for DualContouringMesher { fn default() -> Self { Self { voxel_size: 1.0 } } }
impl DualContouringMesher {
    pub fn with_voxel_size(mut self, voxel_size: f32) -> Self { self.voxel_size = voxel_size; self }
    pub fn mesh(&mut self, volume: &Volume) -> Option<Vec<Vec3f>> {
        let _pair = EXTRACT.lock().unwrap_or_else(|e| e.into_inner());
        let (mut d, mut n) = (ptr::null(), 0usize);
        check(unsafe { ffi::bs_mesh_dc_device(volume.h, self.voxel_size, &mut d, &mut n) });  // panics where the reference panics
        Some(take_vertices(n))
    }
}

/// voxel::meshing::ActiveVoxelsMesher (src/voxel/meshing/active_voxels.rs:4-22)
#[derive(Default)]
pub struct ActiveVoxelsMesher;
impl ActiveVoxelsMesher {
    pub fn mesh(&mut self, volume: &Volume) -> Vec<Vector3<isize>> {
        let (mut p, mut n) = (ptr::null_mut(), 0usize);
        check(unsafe { ffi::bs_mesh_active_voxels(volume.h, &mut p, &mut n) });
        let s = unsafe { std::slice::from_raw_parts(p, n * 3) };
        let out = s.chunks_exact(3).map(|c| Vector3::new(c[0] as isize, c[1] as isize, c[2] as isize)).collect();
        unsafe { ffi::bs_buffer_free(p as *mut _) };
        out
    }
}

/// algo::merge_points (src/algo/merge_points.rs:4-41) for 3-D f32 points
pub struct IndexedVertices { pub points: Vec<Vec3f>, pub indices: Vec<usize> }
pub fn merge_points(points: impl Iterator<Item = Vec3f>) -> IndexedVertices {
    let flat: Vec<f32> = points.flat_map(|p| [p.x, p.y, p.z]).collect();
    let (mut u, mut nu, mut idx) = (ptr::null_mut(), 0usize, ptr::null_mut());
    check(unsafe { ffi::bs_merge_points(ctx(), flat.as_ptr(), flat.len() / 3, &mut u, &mut nu, &mut idx) });
    let points = unsafe { std::slice::from_raw_parts(u, nu * 3) }.chunks_exact(3).map(|c| Vec3f::new(c[0], c[1], c[2])).collect();
    let indices = unsafe { std::slice::from_raw_parts(idx, flat.len() / 3) }.iter().map(|&i| i as usize).collect();
    unsafe { ffi::bs_buffer_free(u as *mut _); ffi::bs_buffer_free(idx as *mut _) };
    IndexedVertices { points, indices }
}

/// io::stl (src/io/stl.rs): the reader leaves the triangles on the device, where `MeshToVolume::convert_device` takes them
pub struct DeviceTriangles { d: *mut f32, n: usize }
impl Drop for DeviceTriangles { fn drop(&mut self) { unsafe { ffi::bs_device_free(ctx(), self.d as *mut _) } } }
pub fn read_stl_to_device(bytes: &[u8]) -> std::io::Result<DeviceTriangles> {
    let (mut d, mut n) = (ptr::null_mut(), 0usize);
    let st = unsafe { ffi::bs_stl_decode(ctx(), bytes.as_ptr(), bytes.len(), &mut d, &mut n) };
    if st == ffi::BS_ERR_INVALID { return Err(std::io::Error::new(std::io::ErrorKind::UnexpectedEof, "short STL buffer")); }  // read_exact in io/stl.rs:47-59
    check(st);
    Ok(DeviceTriangles { d, n })
}
impl MeshToVolume {
    pub fn convert_device(&mut self, tris: &DeviceTriangles) -> Option<Volume> {
        let mut h = ptr::null_mut();
        let st = unsafe { ffi::bs_mesh_to_volume_device(ctx(), tris.d, tris.n, self.voxel_size, self.band_width as i64, &mut h) };
        if st == ffi::BS_ERR_EMPTY_MESH { return None; }
        check(st);
        Some(Volume { h })
    }
}

/// remeshing::voxel::{MeshingMethod, VoxelRemesher} (src/remeshing/voxel.rs:10-95): one FFI call from the triangle soup to
/// the vertex soup; the library converts and extracts slab by slab and overlaps the read-back with the kernels.
#[derive(Clone, Copy, PartialEq, Eq)]
pub enum MeshingMethod { FeaturePreserving, Manifold }
pub struct VoxelRemesher { voxel_size: f32, method: MeshingMethod }
impl Default for VoxelRemesher { fn default() -> Self { Self { voxel_size: 1.0, method: MeshingMethod::Manifold } } }
impl VoxelRemesher {
    pub fn with_voxel_size(mut self, size: f32) -> Self { self.voxel_size = size; self }
    pub fn with_meshing_method(mut self, method: MeshingMethod) -> Self { self.method = method; self }
    /// `remesh` of the reference takes any `TriangleMesh`; the shim's caller hands over its triangle soup (9 floats each).
    pub fn remesh(&mut self, soup: &[Vec3f]) -> Option<Vec<Vec3f>> {
        let n_tris = soup.len() / 3;
        let mut out: Vec<Vec3f> = Vec::with_capacity(std::cmp::max(1024, 12 * n_tris));
        for _ in 0..2 {
            let mut n = 0usize;
            let st = unsafe { ffi::bs_voxel_remesh_into(ctx(), soup.as_ptr() as *const f32, n_tris, self.voxel_size,
                if self.method == MeshingMethod::FeaturePreserving { 1 } else { 0 }, 0, out.as_mut_ptr() as *mut f32, out.capacity() * 3, &mut n) };
            if st == ffi::BS_ERR_EMPTY_MESH { return None; }
            if st == ffi::BS_ERR_INVALID && n > out.capacity() * 3 { out = Vec::with_capacity(n / 3); continue; }  // estimate too small: retry with the reported size
            check(st);
            unsafe { out.set_len(n / 3) };
            return Some(out);
        }
        unreachable!()
    }
}

pub mod prelude {
    pub use super::{DualContouringMesher, MarchingCubesMesher, MeshToVolume, Volume, VolumeBuilder};
}
