// Shared device/host declarations of libbshark_cuda (sm_100a).
//
// Data layout in HBM: a volume is a flat, sorted list of 8^3 bricks
//   keys[n]        u64   brick key, sorted ascending == the reference's leaf visit order
//   values[n*512]  f32   brick-major, inside a brick the reference's leaf layout x<<6 | y<<3 | z
//                        (leaf_node/mod.rs:29-36) so a z-line is one 32 B sector and a brick 2 KB
//   masks[n*8]     u64   512 active bits per brick, bit (o&63) of word (o>>6)
// plus the (rare, CSG-only) active tiles as two small sorted lists.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <map>
#include <mutex>
#include <unordered_set>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../include/bshark.h"

// ---- brick keys -------------------------------------------------------------------------------------
// Brick coordinate b = voxel >> 3 per axis, biased by 2^17 into 18 bits u. The reference's tree is
// Root(BTreeMap keyed by 4096-voxel origin, lexicographic x,y,z) -> 32^3 node -> 16^3 node -> leaf, visited
// in ascending slot offset (x-major). Packing [root x,y,z : 9 bits each][node5 slot x,y,z : 5 each]
// [node4 slot x,y,z : 4 each] makes "ascending key" == that visit order.
#define BS_COORD_BIAS (1 << 17)
#define BS_BRICK_MIN (-(1 << 17))
#define BS_BRICK_MAX ((1 << 17) - 1)
#define BS_KEY_INVALID 0xFFFFFFFFFFFFFFFFull

__host__ __device__ __forceinline__ uint64_t bs_brick_key(int bx, int by, int bz) {
    uint32_t ux = (uint32_t)(bx + BS_COORD_BIAS), uy = (uint32_t)(by + BS_COORD_BIAS), uz = (uint32_t)(bz + BS_COORD_BIAS);
    uint64_t k = 0;
    k |= (uint64_t)(ux >> 9) << 45; k |= (uint64_t)(uy >> 9) << 36; k |= (uint64_t)(uz >> 9) << 27;
    k |= (uint64_t)((ux >> 4) & 31) << 22; k |= (uint64_t)((uy >> 4) & 31) << 17; k |= (uint64_t)((uz >> 4) & 31) << 12;
    k |= (uint64_t)(ux & 15) << 8; k |= (uint64_t)(uy & 15) << 4; k |= (uint64_t)(uz & 15);
    return k;
}
__host__ __device__ __forceinline__ void bs_key_brick(uint64_t k, int& bx, int& by, int& bz) {
    uint32_t ux = (uint32_t)(((k >> 45) & 511) << 9 | ((k >> 22) & 31) << 4 | ((k >> 8) & 15));
    uint32_t uy = (uint32_t)(((k >> 36) & 511) << 9 | ((k >> 17) & 31) << 4 | ((k >> 4) & 15));
    uint32_t uz = (uint32_t)(((k >> 27) & 511) << 9 | ((k >> 12) & 31) << 4 | (k & 15));
    bx = (int)ux - BS_COORD_BIAS; by = (int)uy - BS_COORD_BIAS; bz = (int)uz - BS_COORD_BIAS;
}
// key of the 16^3-brick node (128^3 voxels) and of the 32^3 node (4096^3 voxels) a brick lives in
__host__ __device__ __forceinline__ uint64_t bs_key_node4(uint64_t k) { return k >> 12; }
__host__ __device__ __forceinline__ uint64_t bs_key_node5(uint64_t k) { return k >> 27; }

// sentinel for "no distance yet" in the unsigned field: bytes 0x7F -> 3.39e38, below +inf and NaN bit patterns
#define BS_UDF_SENTINEL_BITS 0x7F7F7F7Fu

// ---- exact (never contracted) f32 arithmetic ----------------------------------------------------------
// The reference is Rust: no FMA contraction, IEEE round-to-nearest. Everything that decides topology or
// must be bit-identical (subdivision, boxes, point-triangle distance, MC/DC vertices, fast sweeping) is
// written with these so -fmad cannot fuse it.
#ifdef __CUDACC__
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float xsqrt(float a) { return __fsqrt_rn(a); }
struct f3 { float x, y, z; };
__device__ __forceinline__ f3 xadd(f3 a, f3 b) { return {xadd(a.x, b.x), xadd(a.y, b.y), xadd(a.z, b.z)}; }
__device__ __forceinline__ f3 xsub(f3 a, f3 b) { return {xsub(a.x, b.x), xsub(a.y, b.y), xsub(a.z, b.z)}; }
__device__ __forceinline__ f3 xscale(f3 a, float s) { return {xmul(a.x, s), xmul(a.y, s), xmul(a.z, s)}; }
// nalgebra 3-vector dot: (x*x' + y*y') + z*z'
__device__ __forceinline__ float xdot(f3 a, f3 b) { return xadd(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xmul(a.z, b.z)); }
__device__ __forceinline__ float xnorm2(f3 a) { return xdot(a, a); }
__device__ __forceinline__ f3 xcross(f3 a, f3 b) {
    return {xsub(xmul(a.y, b.z), xmul(a.z, b.y)), xsub(xmul(a.z, b.x), xmul(a.x, b.z)), xsub(xmul(a.x, b.y), xmul(a.y, b.x))};
}
#endif

// ---- host-side objects --------------------------------------------------------------------------------
struct bs_stat { const char* name; double value; };

// process-wide count of this library's own kernel launches (CUB plumbing not included); bench.py reports the
// difference across its timed region as "gpu_launches"
extern unsigned long long g_bs_launches;
static inline void bs_count_launch() { ++g_bs_launches; }

struct bs_volume;
// products of a marching-cubes count that the matching emit consumes (bs_mc.cu)
struct bs_mc_pending {
    const bs_volume* vol = nullptr; float voxel_size = 0.f; size_t n = 0; unsigned long long n_tris = 0;
    unsigned long long* d_desc = nullptr; unsigned* d_counts = nullptr; unsigned long long* d_offsets = nullptr; int* d_nbr = nullptr;
};
struct bs_context {
    // One in-flight call per context: every entry point of the ABI holds this lock for its whole duration (BS_ENTER),
    // so handles may be used from several host threads. The result of a *_device extraction stays valid until the next
    // extraction on the context: callers that pair bs_mesh_mc_device with bs_context_copy_out_verts from several threads
    // must serialise the pair themselves (bs_mesh_mc / bs_mesh_dc do both under one lock).
    std::recursive_mutex mtx;
    std::unordered_set<bs_volume*> live_volumes;  // handles that still point at this context (bs_context_destroy orphans them)
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    // device memory cache (bs_raw_alloc / bs_raw_free): freed blocks are kept by rounded size and handed out again
    // without any driver call; everything runs on `stream`, so reuse is ordered by the stream itself
    std::multimap<size_t, void*> cache_free;
    struct LiveBlock { size_t size; unsigned long long epoch; };
    std::unordered_map<void*, LiveBlock> cache_live;
    // every entry point of the ABI starts a new epoch (bs_op_begin); an epoch in which bs_fail was called has its
    // still-live blocks handed back to the cache when the next one starts: error paths return early and do not free
    unsigned long long epoch = 0, failed_epoch = 0;
    bool fail_pending = false;
    size_t cache_free_bytes = 0, cache_total_bytes = 0;
    std::string err;
    std::vector<bs_stat> stats;
    int8_t* d_mc33 = nullptr;       // MC33 tables blob (mc33_tables.h)
    // device-side error word of the running call (BS_DERR_*): kernels OR bits into it instead of dropping work silently;
    // bs_convert_impl clears it at the start and turns a non-zero word into BS_ERR_RANGE at its final synchronisation
    unsigned* d_err = nullptr;
    float* d_out_verts = nullptr;   // last extraction result left on the device
    size_t out_verts_cap = 0;
    // pipelined remesh (bs_voxel_remesh_into): second result buffer + copy stream so that slab k's read-back overlaps slab
    // k + 1's kernels; the MC33 c-vertex travels from slab to slab through d_mc_carry while mc_chain is set
    float* d_out_alt = nullptr; size_t out_alt_cap = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    float* d_mc_carry = nullptr; bool mc_chain = false;
    bs_mc_pending mc_pending;
    // small read-backs (bs_fetch / bs_sync): page-locked scratch the device writes directly, so that the counters a call needs
    // on the host never queue behind a bulk copy on a copy engine
    unsigned* h_ctrl = nullptr; unsigned* d_ctrl = nullptr; size_t ctrl_used = 0;
    struct Pending { void* dst; size_t off, bytes; };
    std::vector<Pending> pending;
    // stage timing
    std::vector<std::pair<const char*, cudaEvent_t>> marks;
    // BS_FLAG_COUNT_WORK: instrumented winding-number traversal (node visits, far evals, exact triangles, voxels)
    int count_work = 0;
    // BS_FLAG_SIGN_PROPAGATION (default 1): closed meshes take one winding-number traversal per connected band component
    int sign_propagation = 1;
    // per-convert: is the mesh a closed 2-cycle (bs_signprop.cu)
    bool mesh_closed = false;
    double fwn_counts[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // lane visits, far evals, exact tris, voxels, warp-level visits, traversals, brick-level visits, hoisted nodes
};

struct bs_volume {
    bs_context* ctx = nullptr;
    float voxel_size = 1.0f;
    size_t n_bricks = 0;
    unsigned long long* keys = nullptr;
    float* values = nullptr;
    unsigned long long* masks = nullptr;
    // active tiles: 8^3 tiles keyed like bricks; 128^3 tiles keyed by bs_key_node4 form. Values are +-FLT_MAX.
    size_t n_tiles8 = 0, n_tiles128 = 0;
    unsigned long long* tile8_keys = nullptr; float* tile8_values = nullptr;
    unsigned long long* tile128_keys = nullptr; float* tile128_values = nullptr;
    // multi-GPU (bs_mesh_to_volume_sharded): owned[b] = 1 for this rank's bricks, 0 for read-only halo copies;
    // nullptr = everything owned. Extraction emits owned bricks only.
    unsigned char* owned = nullptr;
    size_t n_owned = 0;
};

#define BS_DERR_STACK 1u      /* winding-number traversal stack full (bs_fwn.cu STACK) */
#define BS_DERR_PROBE 2u      /* brick hash: key not found within the probe limit */

// error plumbing ----------------------------------------------------------------------------------------
bs_status bs_fail(bs_context* ctx, bs_status st, const char* fmt, ...);
#define BS_CUDA(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return bs_fail((ctx), BS_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while (0)
#define BS_TRY(call) do { bs_status s_ = (call); if (s_ != BS_OK) return s_; } while (0)

// device allocations of the library: served from the context's block cache (a steady workload makes no driver calls)
bs_status bs_raw_alloc(bs_context* ctx, size_t bytes, void** out);
void bs_raw_free(bs_context* ctx, void* p);
void bs_cache_release(bs_context* ctx);  // give every cached (free) block back to the driver
void bs_op_begin(bs_context* ctx);       // first thing an ABI entry point does once its handles are validated
// lock the context for the rest of the enclosing scope, select its device, start a new allocation epoch
#define BS_ENTER(c) std::lock_guard<std::recursive_mutex> bs_lock_((c)->mtx); cudaSetDevice((c)->device); bs_op_begin(c)
template <class T> bs_status bs_alloc(bs_context* ctx, T** p, size_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    void* q = nullptr;
    BS_TRY(bs_raw_alloc(ctx, count * sizeof(T), &q));
    *p = (T*)q;
    return BS_OK;
}
template <class T> void bs_free(bs_context* ctx, T* p) { if (p) bs_raw_free(ctx, (void*)p); }

// Read a few words (<= 16 KB, 4-byte granular) of device memory into a host variable WITHOUT a copy engine: a tiny kernel stores them
// into mapped page-locked scratch and bs_sync -- cudaStreamSynchronize on the context's stream -- delivers them to `host_dst`
// (which must stay alive until then). With cudaMemcpyAsync these reads wait behind whatever bulk device-to-host copy another stream has
// queued (measured: the slab read-back of bs_voxel_remesh_into delayed every control read of the next slab by milliseconds).
bs_status bs_fetch(bs_context* ctx, void* host_dst, const void* d_src, size_t bytes);
bs_status bs_sync(bs_context* ctx);
void bs_mark(bs_context* ctx, const char* name);          // record an event named `name` on the stream
void bs_marks_begin(bs_context* ctx);                     // clear marks + stats, record "begin"
void bs_marks_end(bs_context* ctx);                       // sync, turn consecutive marks into "<name>_ms" stats
void bs_stat_add(bs_context* ctx, const char* name, double v);

bs_volume* bs_volume_new(bs_context* ctx, float voxel_size);
bs_status bs_volume_alloc_bricks(bs_volume* v, size_t n);  // keys / values / masks for n bricks

inline unsigned bs_blocks(size_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

// stage entry points (one .cu each)
// what the slabs of one mesh share (bs_voxel_remesh_into converts a mesh slab by slab): the closedness verdict and the
// coarse cut of the key space are computed by the first slab and reused by the others
struct bs_convert_plan { bool valid = false; bool closed = false; int world = 0; std::vector<unsigned long long> bounds; };
bs_status bs_convert_impl(bs_context* ctx, const float* d_tris, size_t n_tris, float voxel_size, int64_t band,
                          int rank, int world, bs_volume** out, bs_convert_plan* plan = nullptr);
bs_status bs_sign_impl(bs_context* ctx, const float* d_tris, size_t n_tris, bs_volume* vol, const unsigned long long* d_touches /*per brick, may be null*/,
                       const unsigned long long* d_blk /*blocked lattice edges, 24 words per brick; null = per-voxel signs*/);
bs_status bs_mc_impl(const bs_volume* v, float voxel_size, const float** d_verts, size_t* n_verts);
bs_status bs_mc_count_phase(const bs_volume* v, float voxel_size, size_t* n_verts);  // volumes without active tiles
bs_status bs_mc_emit_phase(const bs_volume* v, float* const* dst, int world, size_t offset_floats, size_t cap_floats);
void bs_mc_pending_release(bs_context* ctx);
bs_status bs_dc_impl(const bs_volume* v, float voxel_size, const float** d_verts, size_t* n_verts);
bs_status bs_csg_impl(bs_volume* a, bs_volume* b, int op, bs_volume** out);
bs_status bs_offset_impl(bs_volume* a, float distance, bs_volume** out);
bs_status bs_from_voxels_impl(bs_context* ctx, const int32_t* d_ijk, const float* d_values, size_t m, float voxel_size,
                              bs_volume** out);
bs_status bs_stl_decode_impl(bs_context* ctx, const unsigned char* d_stl, size_t n_bytes, float** d_tris, size_t* n_tris);
bs_status bs_stl_encode_impl(bs_context* ctx, const float* d_verts, size_t n_verts, unsigned char** d_stl, size_t* n_bytes);
bs_status bs_active_voxels_impl(const bs_volume* v, int** d_verts, size_t* n_verts);
bs_status bs_merge_points_impl(bs_context* ctx, const float* d_pts, size_t n, float** d_unique, size_t* n_unique, unsigned** d_indices);
// sign propagation on closed meshes (bs_signprop.cu)
struct bs_closed_check { bool exact, closed, pending; unsigned long long* d_sums; int* d_bad; unsigned long long h_sums[4]; int h_bad; };
bs_status bs_mesh_closed_begin(bs_context* ctx, const float* d_tris, size_t n_tris, bs_closed_check* chk);  // enqueue; verdict after the next stream sync
bool bs_mesh_closed_finish(bs_context* ctx, bs_closed_check* chk);
struct bs_sign_components {  // per-brick component tables (8 components per brick)
    bool ok;
    unsigned long long *comp /*[n][8][8] masks*/, *planes /*[n][8][6] face planes*/, *face /*[n][9]*/, *rest /*[n][8]*/, *seed /*[n][8] voxels to evaluate*/;
    unsigned char* ncomp; unsigned short* first /*[n][8] first voxel of a component*/; unsigned* par /*[n][8] union-find over (brick, component)*/;
};
bs_status bs_sign_components_impl(bs_context* ctx, const bs_volume* vol, const unsigned long long* d_blk, bs_sign_components* C, unsigned* d_nchunks, int per_chunk, unsigned long long* d_nseeds);
void bs_sign_components_free(bs_context* ctx, bs_sign_components* C);
int bs_sign_brute_max();
bs_status bs_sign_brute_impl(bs_context* ctx, const float* d_tris, size_t n_tris, bs_volume* vol, const bs_sign_components* C, unsigned n_seeds);
bs_status bs_sign_broadcast_impl(bs_context* ctx, bs_volume* vol, const bs_sign_components* C);
bs_status bs_builder_impl(bs_context* ctx, int kind, float voxel_size, const float* p, bs_volume** out);
