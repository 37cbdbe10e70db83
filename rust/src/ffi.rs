//! Raw bindings to include/bshark.h (the only place the crate touches CUDA).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_float, c_int, c_void};

#[repr(C)] pub struct bs_context { _p: [u8; 0] }
#[repr(C)] pub struct bs_volume { _p: [u8; 0] }
pub type bs_status = c_int;
pub const BS_OK: bs_status = 0;
pub const BS_ERR_EMPTY_MESH: bs_status = 1;
pub const BS_ERR_INVALID: bs_status = 3;
pub const BS_ERR_REFERENCE_PANICS: bs_status = 4;

extern "C" {
    pub fn bs_context_create(device: c_int, out: *mut *mut bs_context) -> bs_status;
    pub fn bs_context_destroy(ctx: *mut bs_context);
    pub fn bs_last_error(ctx: *const bs_context) -> *const c_char;
    pub fn bs_mesh_to_volume(ctx: *mut bs_context, tris: *const c_float, n_tris: usize, voxel_size: c_float, band_width: i64, out: *mut *mut bs_volume) -> bs_status;
    pub fn bs_volume_from_voxels(ctx: *mut bs_context, ijk: *const i32, values: *const c_float, m: usize, voxel_size: c_float, out: *mut *mut bs_volume) -> bs_status;
    pub fn bs_volume_empty(ctx: *mut bs_context, voxel_size: c_float, out: *mut *mut bs_volume) -> bs_status;
    pub fn bs_volume_sphere(ctx: *mut bs_context, voxel_size: c_float, radius: c_float, origin: *const c_float, out: *mut *mut bs_volume) -> bs_status;
    pub fn bs_volume_cuboid(ctx: *mut bs_context, voxel_size: c_float, min: *const c_float, max: *const c_float, out: *mut *mut bs_volume) -> bs_status;
    pub fn bs_volume_iwp(ctx: *mut bs_context, voxel_size: c_float, min: *const c_float, max: *const c_float, cell_size: c_float, out: *mut *mut bs_volume) -> bs_status;
    pub fn bs_volume_clone(v: *const bs_volume, out: *mut *mut bs_volume) -> bs_status;
    pub fn bs_volume_free(v: *mut bs_volume);
    pub fn bs_volume_voxel_size(v: *const bs_volume) -> c_float;
    pub fn bs_volume_union(a: *mut bs_volume, b: *mut bs_volume, out: *mut *mut bs_volume) -> bs_status;
    pub fn bs_volume_subtract(a: *mut bs_volume, b: *mut bs_volume, out: *mut *mut bs_volume) -> bs_status;
    pub fn bs_volume_intersect(a: *mut bs_volume, b: *mut bs_volume, out: *mut *mut bs_volume) -> bs_status;
    pub fn bs_volume_offset(a: *mut bs_volume, distance: c_float, out: *mut *mut bs_volume) -> bs_status;
    pub fn bs_mesh_mc_device(v: *const bs_volume, voxel_size: c_float, d_verts: *mut *const c_float, n_verts: *mut usize) -> bs_status;
    pub fn bs_mesh_dc_device(v: *const bs_volume, voxel_size: c_float, d_verts: *mut *const c_float, n_verts: *mut usize) -> bs_status;
    pub fn bs_context_copy_out_verts(ctx: *mut bs_context, dst: *mut c_float, n_floats: usize) -> bs_status;
    pub fn bs_buffer_free(p: *mut c_void);
    pub fn bs_device_free(ctx: *mut bs_context, d_ptr: *mut c_void);
    pub fn bs_mesh_to_volume_device(ctx: *mut bs_context, d_tris: *const c_float, n_tris: usize, voxel_size: c_float, band_width: i64, out: *mut *mut bs_volume) -> bs_status;
    pub fn bs_stl_decode(ctx: *mut bs_context, stl: *const u8, n_bytes: usize, d_tris: *mut *mut c_float, n_tris: *mut usize) -> bs_status;
    pub fn bs_stl_encode(ctx: *mut bs_context, d_verts: *const c_float, n_verts: usize, stl: *mut *mut u8, n_bytes: *mut usize) -> bs_status;
    pub fn bs_mesh_active_voxels(v: *const bs_volume, verts: *mut *mut i32, n_verts: *mut usize) -> bs_status;
    pub fn bs_merge_points(ctx: *mut bs_context, points: *const c_float, n: usize, unique: *mut *mut c_float, n_unique: *mut usize, indices: *mut *mut u32) -> bs_status;
}
