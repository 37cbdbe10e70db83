/* bshark.h -- C ABI of libbshark_cuda.so: the B200 (sm_100a) implementation of baby_shark's
 * implicit-modelling hot path (mesh -> narrow-band SDF -> CSG / offset -> iso-surface extraction).
 *
 * The reference (Rust crate baby_shark 0.3.12) has no FFI for this path; these entry points are what a
 * drop-in `voxel` module binds instead of its CPU implementation.  Each one cites the reference
 * interface it replaces (paths under the reference checkout).  INTEGRATION.md shows the Rust shim.
 *
 * Conventions
 *  - Opaque handles own device memory.  `bs_volume` lives on the device between calls; host<->device
 *    copies happen only for triangles in (bs_mesh_to_volume) and vertices out (bs_mesh_mc / bs_mesh_dc).
 *  - Consuming operations mirror Rust move semantics: inputs marked "consumed" are freed by the call,
 *    whether it succeeds or not, and must not be used again.
 *  - Every call is synchronous at return and never throws or unwinds across the ABI.
 *  - `bs_status` 0 = ok.  The shim maps BS_ERR_EMPTY_MESH to `None` (reference returns None) and
 *    BS_ERR_REFERENCE_PANICS to a panic (the reference `todo!()` / `unwrap()` / `unreachable!()`s there).
 *  - There is NO CPU fallback: without a CUDA device every entry point returns BS_ERR_NO_DEVICE.
 *  - Voxel indices must lie in [-2^20, 2^20) per axis (BS_ERR_RANGE otherwise); the reference is
 *    unbounded (isize indices, BTreeMap root).
 *  - Every entry point locks its context for the duration of the call: handles may be shared between host threads.
 *    The result of a *_device extraction lives in a per-context buffer until the next extraction; a caller that pairs
 *    bs_mesh_mc_device with bs_context_copy_out_verts from several threads serialises the pair itself (bs_mesh_mc /
 *    bs_mesh_dc do both under one lock).
 *  - bs_context_destroy with volumes still alive orphans them: bs_volume_free on such a handle is safe, everything else
 *    is BS_ERR_INVALID.
 */
#ifndef BSHARK_H
#define BSHARK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bs_context bs_context;
typedef struct bs_volume bs_volume;
typedef int bs_status;

enum {
    BS_OK = 0,
    BS_ERR_EMPTY_MESH = 1,       /* MeshToVolume::convert -> None (mesh_to_volume.rs:58-60) */
    BS_ERR_CUDA = 2,             /* CUDA runtime error; see bs_last_error */
    BS_ERR_INVALID = 3,          /* null handle / bad argument / volumes from different contexts */
    BS_ERR_REFERENCE_PANICS = 4, /* input on which the reference panics (tiles in DC / offset, ...) */
    BS_ERR_RANGE = 5,            /* voxel index outside [-2^20, 2^20) */
    BS_ERR_NO_DEVICE = 6,        /* no CUDA device / not an sm_100 device: there is no CPU fallback */
    BS_ERR_UNSUPPORTED = 7       /* defined in the reference but not implemented on the device yet */
};

/* ---- context ------------------------------------------------------------------------------------- */
/* One CUDA device, one stream, one stream-ordered memory pool.  device = -1 uses the current device. */
bs_status bs_context_create(int device, bs_context** out);
void bs_context_destroy(bs_context* ctx);
const char* bs_last_error(const bs_context* ctx);
int bs_context_device(const bs_context* ctx);
/* The context's stream as a cudaStream_t (void* to keep CUDA headers out of the ABI). */
void* bs_context_stream(const bs_context* ctx);

/* ---- mesh -> volume ------------------------------------------------------------------------------- */
/* Replaces MeshToVolume::convert (src/voxel/mesh_to_volume.rs:52-73) with voxel_size / band_width from
 * with_voxel_size / with_narrow_band_width (:28-50).  `tris` = n x 9 host floats (p1,p2,p3 per triangle,
 * the order `Triangles::triangles()` yields, src/mesh/traits.rs:4-8). */
bs_status bs_mesh_to_volume(bs_context* ctx, const float* tris, size_t n_tris, float voxel_size,
                            int64_t band_width, bs_volume** out);
/* Same, triangles already resident on the context's device (n x 9 floats). */
bs_status bs_mesh_to_volume_device(bs_context* ctx, const float* d_tris, size_t n_tris, float voxel_size,
                                   int64_t band_width, bs_volume** out);
/* Brick-sharded variant for multi-GPU runs: the mesh is replicated, rank r of `world` keeps the r-th contiguous
 * slab of the brick list (in the reference's leaf visit order) plus, as read-only halo, the +x/+y/+z neighbour
 * bricks its cells need. Extraction on the result emits only the cells of owned bricks, so concatenating the
 * per-rank outputs in rank order gives exactly the single-GPU output (DESIGN.md "Multi-GPU"). */
bs_status bs_mesh_to_volume_sharded(bs_context* ctx, const float* d_tris, size_t n_tris, float voxel_size,
                                    int64_t band_width, int rank, int world, bs_volume** out);

/* ---- volume --------------------------------------------------------------------------------------- */
/* Volume::from_fn backing (src/voxel/volume/mod.rs:40-72): the shim evaluates the closure on the host
 * exactly as :53-67 and passes the kept voxels (m x 3 indices, m values). */
bs_status bs_volume_from_voxels(bs_context* ctx, const int32_t* ijk, const float* values, size_t m,
                                float voxel_size, bs_volume** out);
/* Volume::with_voxel_size (volume/mod.rs:18-24): empty volume. */
bs_status bs_volume_empty(bs_context* ctx, float voxel_size, bs_volume** out);
/* VolumeBuilder::{sphere,cuboid,iwp} (src/voxel/volume/builder.rs:21-76), evaluated on the device. */
bs_status bs_volume_sphere(bs_context* ctx, float voxel_size, float radius, const float origin[3], bs_volume** out);
bs_status bs_volume_cuboid(bs_context* ctx, float voxel_size, const float min[3], const float max[3], bs_volume** out);
bs_status bs_volume_iwp(bs_context* ctx, float voxel_size, const float min[3], const float max[3], float cell_size,
                        bs_volume** out);
/* impl Clone for Volume (volume/mod.rs:116-123) */
bs_status bs_volume_clone(const bs_volume* v, bs_volume** out);
void bs_volume_free(bs_volume* v);
/* Volume::voxel_size (volume/mod.rs:31-34) */
float bs_volume_voxel_size(const bs_volume* v);

/* Volume::{union,subtract,intersect}(self, other) -> Self (volume/mod.rs:74-93): flood-fill both, then CSG.
 * Both inputs are consumed. */
bs_status bs_volume_union(bs_volume* self_consumed, bs_volume* other_consumed, bs_volume** out);
bs_status bs_volume_subtract(bs_volume* self_consumed, bs_volume* other_consumed, bs_volume** out);
bs_status bs_volume_intersect(bs_volume* self_consumed, bs_volume* other_consumed, bs_volume** out);
/* Volume::offset(self, distance) -> Self (volume/mod.rs:95-108). Input consumed. */
bs_status bs_volume_offset(bs_volume* self_consumed, float distance, bs_volume** out);

/* ---- extraction ----------------------------------------------------------------------------------- */
/* MarchingCubesMesher::mesh (src/voxel/meshing/marching_cubes.rs:43-63) with with_voxel_size (:32-36):
 * returns the vertex soup, 3 consecutive xyz per triangle, in the reference's emission order.
 * *verts is library-owned host memory (bs_buffer_free). */
bs_status bs_mesh_mc(const bs_volume* v, float voxel_size, float** verts, size_t* n_verts);
/* DualContouringMesher::mesh (src/voxel/meshing/dual_contouring.rs:23-83). The reference's output order is
 * nondeterministic (rayon + Mutex); this returns leaves in visit order. BS_ERR_REFERENCE_PANICS where the
 * reference hits todo!() (active tiles) or unreachable!() (:340). */
bs_status bs_mesh_dc(const bs_volume* v, float voxel_size, float** verts, size_t* n_verts);
/* MarchingCubesMesher::mesh (marching_cubes.rs:43-63) in two steps, for the multi-GPU output exchange: the count returns the
 * number of vertices this volume (this rank's slab of a sharded volume) will emit; once the ranks have exchanged their counts,
 * the emit writes the triangles at float offset `offset_floats` of EVERY buffer in dst[0..world) -- the ranks' result buffers,
 * mapped with bs_ipc_open -- so the exchange over NVLink rides on the emission, tile by tile, instead of following it (every
 * buffer must hold cap_floats floats). The emit is asynchronous on the context's stream with respect to the peers: the caller
 * fences (bench.py: a one-element all-reduce on the same stream). Volumes with active tiles: BS_ERR_UNSUPPORTED. */
bs_status bs_mesh_mc_count(const bs_volume* v, float voxel_size, size_t* n_verts);
bs_status bs_mesh_mc_emit_push(const bs_volume* v, float* const* dst, int world, size_t offset_floats, size_t cap_floats);
/* VoxelRemesher::remesh (src/remeshing/voxel.rs:64-84): MeshToVolume::convert with band width 0, then
 * MarchingCubesMesher (method 0 = MeshingMethod::Manifold, the default) or DualContouringMesher (method 1 =
 * FeaturePreserving), in ONE call from host triangles (9 floats each) to the vertex soup in caller-owned host memory
 * (`dst`, room for cap_floats floats; page-locked memory lets the read-back overlap the kernels). The mesh is converted
 * and extracted in `slabs` contiguous pieces of the reference's leaf visit order (0 = pick: 4 from ~1 M triangles, else
 * 1; env BSHARK_REMESH_SLABS), the copy of one piece running while the next is computed; the result is the same vertex
 * array, bit for bit, as bs_mesh_to_volume + bs_mesh_mc / bs_mesh_dc. *n_floats receives the size of the result; when it
 * exceeds cap_floats the call returns BS_ERR_INVALID and the caller retries with a buffer of that size.
 * BS_ERR_EMPTY_MESH where the reference returns None. */
bs_status bs_voxel_remesh_into(bs_context* ctx, const float* tris, size_t n_tris, float voxel_size, int method, int slabs,
                               float* dst, size_t cap_floats, size_t* n_floats);
/* Same, leaving the vertices on the device: *d_verts is a device pointer (n_verts x 3 floats) owned by the
 * context and valid until the next extraction call on the same context or bs_context_destroy. */
bs_status bs_mesh_mc_device(const bs_volume* v, float voxel_size, const float** d_verts, size_t* n_verts);
bs_status bs_mesh_dc_device(const bs_volume* v, float voxel_size, const float** d_verts, size_t* n_verts);
void bs_buffer_free(void* p);

/* ---- data formats either side of the path ---------------------------------------------------------------- */
/* StlReader::read_from_buffer (src/io/stl.rs:65-95): binary STL bytes (80-byte header, u32 count, 50-byte records)
 * -> count x 9 floats in device memory, ready for bs_mesh_to_volume_device. *d_tris is library-owned device memory
 * (bs_device_free). A buffer shorter than its header announces is BS_ERR_INVALID (the reference: ReadError). */
bs_status bs_stl_decode(bs_context* ctx, const unsigned char* stl, size_t n_bytes, float** d_tris, size_t* n_tris);
bs_status bs_stl_decode_device(bs_context* ctx, const unsigned char* d_stl, size_t n_bytes, float** d_tris,
                               size_t* n_tris);
/* StlWriter::write_to_buffer (src/io/stl.rs:143-191) for a vertex soup on the device (e.g. the result of
 * bs_mesh_mc_device): zero header, count, per face the recomputed normal (zeros if degenerate), the three vertices
 * and a zero attribute. *stl is library-owned host memory (bs_buffer_free), *d_stl device memory (bs_device_free). */
bs_status bs_stl_encode(bs_context* ctx, const float* d_verts, size_t n_verts, unsigned char** stl, size_t* n_bytes);
bs_status bs_stl_encode_device(bs_context* ctx, const float* d_verts, size_t n_verts, unsigned char** d_stl,
                               size_t* n_bytes);
/* ActiveVoxelsMesher::mesh (src/voxel/meshing/active_voxels.rs:12-22): two triangles per exposed voxel face, integer
 * vertices (the reference returns Vector3<isize>), 3 consecutive xyz per triangle, in the reference's order.
 * Active tiles contribute their boundary voxels (duplicates included, :128-151). Host result: bs_buffer_free; device
 * result: bs_device_free. */
bs_status bs_mesh_active_voxels(const bs_volume* v, int32_t** verts, size_t* n_verts);
bs_status bs_mesh_active_voxels_device(const bs_volume* v, int32_t** d_verts, size_t* n_verts);
/* merge_points (src/algo/merge_points.rs:12-41): exactly coincident points share an index; unique points keep
 * first-occurrence order. unique = n_unique x 3 floats, indices = n entries. */
bs_status bs_merge_points(bs_context* ctx, const float* points, size_t n, float** unique, size_t* n_unique,
                          uint32_t** indices);
bs_status bs_merge_points_device(bs_context* ctx, const float* d_points, size_t n, float** d_unique,
                                 size_t* n_unique, uint32_t** d_indices);
/* MarchingCubesMesher::mesh followed by merge_points, without leaving the device: what `T::from_triangles_soup` does
 * for an indexed mesh type (src/remeshing/voxel.rs:73-83 -> mesh/corner_table/builder.rs:294), minus the host hash
 * pass and with half the bytes to read back. points = n_points x 3 floats (first-occurrence order), indices = one per
 * soup vertex (3 per triangle). Host results: bs_buffer_free. */
bs_status bs_mesh_mc_indexed(const bs_volume* v, float voxel_size, float** points, size_t* n_points, uint32_t** indices,
                             size_t* n_indices);
/* Same, results left on the device (bs_device_free). */
bs_status bs_mesh_mc_indexed_device(const bs_volume* v, float voxel_size, float** d_points, size_t* n_points,
                                    uint32_t** d_indices, size_t* n_indices);
/* Copies `bytes` from device memory of the context's device into caller memory (pinned or pageable) on the context's
 * stream and waits: lets the shim fill buffers it owns from any *_device result. */
bs_status bs_copy_to_host(bs_context* ctx, const void* d_src, void* dst, size_t bytes);
/* Frees device memory returned by the *_device entry points above. */
void bs_device_free(bs_context* ctx, void* d_ptr);

/* ---- parity / debug -------------------------------------------------------------------------------- */
/* Leaves in the reference's visit order (root map order, ascending child offsets).  brick_ijk = leaf
 * origins (n x 3), values = n x 512 in the reference's leaf layout (x<<6 | y<<3 | z, leaf_node/mod.rs:29-36),
 * masks = n x 8 words, bit (o & 63) of word (o >> 6) = voxel o active.  Active tiles (only CSG creates them)
 * come back as origin / edge length in voxels / value.  All buffers are library-owned host memory. */
bs_status bs_volume_download(const bs_volume* v, int32_t** brick_ijk, float** values, uint64_t** masks,
                             size_t* n_bricks, int32_t** tile_ijk, int32_t** tile_size, float** tile_values,
                             size_t* n_tiles);
/* counts only: bricks, active voxels, negative active voxels, active tiles */
bs_status bs_volume_counts(const bs_volume* v, size_t* n_bricks, size_t* n_active, size_t* n_negative,
                           size_t* n_tiles);

/* Copy the first n_floats of the last *_device extraction result into caller memory (pinned or pageable): lets
 * the shim fill a Vec<Vec3f> it allocated itself instead of taking a library-owned buffer. */
bs_status bs_context_copy_out_verts(bs_context* ctx, float* dst, size_t n_floats);
/* Same into device memory of the context's device (e.g. the send buffer of the multi-GPU all-gather). */
bs_status bs_context_copy_out_verts_device(bs_context* ctx, float* d_dst, size_t n_floats);

/* ---- multi-GPU output exchange over NVLink peer memory (one process per GPU) ---------------------------------------
 * The brick-sharded remesh ends with every rank holding the whole triangle soup (north_star: "all-gather the compacted
 * output triangle buffers"). Instead of a collective library the ranks write their slices straight into each other's
 * result buffers: bs_ipc_alloc gives a device buffer plus a 64-byte CUDA IPC handle that the host exchanges by any means
 * (bench.py: torch.distributed all_gather_object), bs_ipc_open maps a peer's buffer into this process (peer access over
 * NVLink / NVSwitch is enabled by the mapping), and bs_context_push_out_verts copies the first n_floats of the last *_device
 * extraction result into `world` destinations (own buffer included) at element offset `offset_floats` with ONE kernel --
 * every byte crosses NVLink once, as P2P stores issued by the producing GPU. The call returns when the copies have been
 * issued on the context's stream; the caller orders consumers behind it (a one-element all-reduce on that stream is the
 * "everyone has delivered" fence bench.py uses). BS_ERR_INVALID when there is no extraction result of that size. */
bs_status bs_ipc_alloc(bs_context* ctx, size_t bytes, void** d_ptr, unsigned char handle[64]);
bs_status bs_ipc_open(bs_context* ctx, const unsigned char handle[64], void** d_ptr);
bs_status bs_ipc_close(bs_context* ctx, void* d_ptr);   /* a pointer from bs_ipc_open */
bs_status bs_ipc_free(bs_context* ctx, void* d_ptr);    /* a pointer from bs_ipc_alloc */
bs_status bs_context_push_out_verts(bs_context* ctx, float* const* dst, int world, size_t offset_floats, size_t n_floats);

/* BS_FLAG_COUNT_WORK = 1: the next bs_mesh_to_volume* calls run the instrumented winding-number traversal and
 * report fwn_visits / fwn_far / fwn_exact_tris / fwn_voxels through bs_context_last_stats (roofline work counts;
 * slower, never used in a timed region).
 * BS_FLAG_SIGN_PROPAGATION = 2 (default 1): on a closed input mesh (every directed edge matched by its reverse) the
 * winding number is traversed once per connected component of the band -- certified from the unsigned distances --
 * instead of once per voxel (mesh_to_volume.rs:198-281 evaluates every voxel; same signs). 0 = always
 * per voxel. Open meshes always take the per-voxel path. */
enum { BS_FLAG_COUNT_WORK = 1, BS_FLAG_SIGN_PROPAGATION = 2 };
bs_status bs_context_set_flag(bs_context* ctx, int flag, int value);

/* Number of kernels of this library launched by the calling process so far (all contexts; library sorts / scans
 * not included). One in-flight call per context: read it between calls. */
unsigned long long bs_kernel_launch_count(void);

/* Per-stage device timings (ms, CUDA events on the context stream) and work counters of the most recent
 * call on the context; names are listed in DESIGN.md.  Returns the number of entries written (<= cap). */
size_t bs_context_last_stats(const bs_context* ctx, const char** names, double* values, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* BSHARK_H */
