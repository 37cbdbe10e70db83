// EXPERIMENTAL, OFF BY DEFAULT (enabled by the environment variable BSHARK_SIGN_PROPAGATION; written at the end of
// round 1: its algorithm is checked on the CPU by tools/sign_propagation_probe.py and tools/signprop_emulation.py, and
// with the last GPU seconds of the round test_experimental_sign_propagation_matches_per_voxel_signs passed on a B200
// (bit-identical volumes on three closed test meshes, fallback on an open one). It has NOT been timed or run at the
// benchmark size yet, so the default path never calls it).
//
// Sign propagation for closed meshes. MeshToVolume::compute_sings (mesh_to_volume.rs:198-281)
// evaluates the winding number of every active voxel. On a closed, consistently oriented mesh the winding number is an
// integer that only changes across the surface, so it is the same for two lattice neighbours p, q whenever the segment
// pq cannot meet the surface: min(d_p, vs) + min(d_q, vs) > |pq| (the scatter-min distance of bs_convert.cu is exact
// below one voxel, because a triangle whose integer box misses a lattice point is at least one voxel away from it).
// Steps: (1) the mesh is closed and consistently oriented iff every directed edge occurs once and its reverse once
// (vertices identified by exact coordinates, like merge_points); (2) union-find over the certified face links of the
// band voxels, inside bricks in shared memory, across brick faces in global memory; (3) the per-voxel traversal of
// bs_fwn.cu runs on one voxel per component only (~5-8 % of the band); (4) every other voxel copies the sign of its
// component's representative. Anything irregular (open mesh, repeated vertex in a triangle, too many voxels) falls back
// to the per-voxel path, which is the reference's semantics for such input.
#include "bs_common.cuh"
#include <cub/cub.cuh>

namespace {

typedef unsigned long long u64;
constexpr unsigned SP_EMPTY = 0xFFFFFFFFu;
constexpr float SP_MARGIN = 1.001f;  // a flat surface exactly between p and q gives d_p + d_q = |pq| up to rounding

// ---- (1) closed and consistently oriented? ------------------------------------------------------------------------------
__device__ __forceinline__ unsigned sp_hash(float x, float y, float z) {
    const unsigned a = __float_as_uint(x + 0.0f), b = __float_as_uint(y + 0.0f), c = __float_as_uint(z + 0.0f);
    unsigned h = a * 73856093u ^ b * 19349663u ^ c * 83492791u;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
    return h;
}
// vertex id = slot of the vertex's coordinate class in an open-addressing table (one representative index per slot)
__global__ void k_sp_vertex_ids(const float* __restrict__ pts, size_t n, unsigned* table, unsigned mask, unsigned* vid) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    unsigned s = sp_hash(x, y, z) & mask;
    for (;;) {
        unsigned cur = table[s];
        if (cur == SP_EMPTY) { cur = atomicCAS(table + s, SP_EMPTY, (unsigned)i); if (cur == SP_EMPTY) break; }
        if (pts[3 * (size_t)cur] == x && pts[3 * (size_t)cur + 1] == y && pts[3 * (size_t)cur + 2] == z) break;
        s = (s + 1) & mask;  // (a NaN vertex never matches: it ends up alone, its edges unmatched -> "not closed")
    }
    vid[i] = s;
}
__global__ void k_sp_edges(const unsigned* __restrict__ vid, size_t n_tris, u64* keys, int* bad) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= n_tris) return;
    const unsigned a = vid[3 * t], b = vid[3 * t + 1], c = vid[3 * t + 2];
    if (a == b || b == c || c == a) *bad = 1;  // a triangle with a repeated vertex: be conservative
    keys[3 * t] = ((u64)a << 32) | b; keys[3 * t + 1] = ((u64)b << 32) | c; keys[3 * t + 2] = ((u64)c << 32) | a;
}
__global__ void k_sp_edges_check(const u64* __restrict__ keys /*sorted*/, size_t m, int* bad) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    const u64 k = keys[i];
    if (i + 1 < m && keys[i + 1] == k) { *bad = 1; return; }  // the same directed edge twice: non-manifold or inconsistent
    const u64 rev = (k << 32) | (k >> 32);
    size_t lo = 0, hi = m;
    while (lo < hi) { const size_t mid = (lo + hi) >> 1; if (keys[mid] < rev) lo = mid + 1; else hi = mid; }
    if (!(lo < m && keys[lo] == rev)) *bad = 1;               // boundary edge
}

// ---- (2) union-find over certified links ------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned sp_find(volatile unsigned* par, unsigned x) {
    unsigned p;
    while ((p = par[x]) != x) x = p;
    return x;
}
// lock-free union by index (the larger root is hooked under the smaller one); works on shared or global memory
__device__ __forceinline__ void sp_union(unsigned* par, unsigned a, unsigned b) {
    for (;;) {
        a = sp_find(par, a); b = sp_find(par, b);
        if (a == b) return;
        if (a > b) { const unsigned t = a; a = b; b = t; }
        if (atomicCAS(par + b, b, a) == b) return;
    }
}
// one CTA (512 threads, one per voxel) per brick: components of the brick's own certified links, written as global ids
__global__ void __launch_bounds__(512) k_sp_bricks(const float* __restrict__ values, const u64* __restrict__ masks, float vs, unsigned* par) {
    __shared__ unsigned s_par[512];
    __shared__ float s_cap[512];
    const size_t b = blockIdx.x;
    const unsigned t = threadIdx.x;
    const bool act = (masks[b * 8 + (t >> 6)] >> (t & 63)) & 1ull;
    const float cap = act ? fminf(fabsf(values[b * 512 + t]), vs) : -1.0f;
    s_par[t] = t; s_cap[t] = cap;
    __syncthreads();
    const float thr = vs * SP_MARGIN;
    if (act) {
        const unsigned x = t >> 6, y = (t >> 3) & 7, z = t & 7;
        if (x < 7 && s_cap[t + 64] >= 0.f && cap + s_cap[t + 64] > thr) sp_union(s_par, t, t + 64);
        if (y < 7 && s_cap[t + 8] >= 0.f && cap + s_cap[t + 8] > thr) sp_union(s_par, t, t + 8);
        if (z < 7 && s_cap[t + 1] >= 0.f && cap + s_cap[t + 1] > thr) sp_union(s_par, t, t + 1);
    }
    __syncthreads();
    par[b * 512 + t] = act ? (unsigned)(b * 512) + sp_find(s_par, t) : SP_EMPTY;
}
__device__ __forceinline__ long long sp_find_key(const u64* keys, size_t n, u64 k) {
    size_t lo = 0, hi = n;
    while (lo < hi) { const size_t mid = (lo + hi) >> 1; if (keys[mid] < k) lo = mid + 1; else hi = mid; }
    return (lo < n && keys[lo] == k) ? (long long)lo : -1;
}
// one CTA (192 threads) per brick: certified links across its +x, +y, +z faces
__global__ void __launch_bounds__(192) k_sp_faces(const u64* __restrict__ keys, size_t n, const float* __restrict__ values, const u64* __restrict__ masks, float vs, unsigned* par) {
    __shared__ long long s_nb[3];
    const size_t b = blockIdx.x;
    if (threadIdx.x < 3) {
        int bx, by, bz; bs_key_brick(keys[b], bx, by, bz);
        const int ax = threadIdx.x;
        bx += ax == 0; by += ax == 1; bz += ax == 2;
        s_nb[ax] = (bx <= BS_BRICK_MAX && by <= BS_BRICK_MAX && bz <= BS_BRICK_MAX) ? sp_find_key(keys, n, bs_brick_key(bx, by, bz)) : -1;
    }
    __syncthreads();
    const unsigned ax = threadIdx.x >> 6, u = (threadIdx.x >> 3) & 7, v = threadIdx.x & 7;
    const long long nb = s_nb[ax];
    if (nb < 0) return;
    const unsigned op = ax == 0 ? ((7u << 6) | (u << 3) | v) : (ax == 1 ? ((u << 6) | (7u << 3) | v) : ((u << 6) | (v << 3) | 7u));
    const unsigned oq = ax == 0 ? ((u << 3) | v) : (ax == 1 ? ((u << 6) | v) : ((u << 6) | (v << 3)));
    const bool ap = (masks[b * 8 + (op >> 6)] >> (op & 63)) & 1ull, aq = (masks[(size_t)nb * 8 + (oq >> 6)] >> (oq & 63)) & 1ull;
    if (!ap || !aq) return;
    const float cp = fminf(fabsf(values[b * 512 + op]), vs), cq = fminf(fabsf(values[(size_t)nb * 512 + oq]), vs);
    if (cp + cq > vs * SP_MARGIN) sp_union(par, (unsigned)(b * 512 + op), (unsigned)((size_t)nb * 512 + oq));
}
// flatten + seed masks: a voxel is its component's representative iff it is its own root
__global__ void __launch_bounds__(512) k_sp_flatten(unsigned* par, u64* seed_masks) {
    const size_t b = blockIdx.x;
    const unsigned t = threadIdx.x;
    const unsigned g = (unsigned)(b * 512 + t);
    bool seed = false;
    if (par[g] != SP_EMPTY) { const unsigned r = sp_find(par, g); par[g] = r; seed = r == g; }  // roots are final: compressing while others read is safe
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, seed);
    if ((t & 31) == 0) reinterpret_cast<unsigned*>(seed_masks + b * 8)[t >> 5] = bal;  // bit t of the brick's 512-bit mask
}
__global__ void k_sp_chunks(const u64* __restrict__ seed_masks, size_t n_bricks, unsigned* n_chunks, int per_chunk) {
    const size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (b >= n_bricks) return;
    int c = 0;
    for (int k = 0; k < 8; ++k) c += __popcll(seed_masks[b * 8 + k]);
    n_chunks[b] = (unsigned)((c + per_chunk - 1) / per_chunk);
}
// (4) every non-representative voxel takes the sign its representative got from the traversal
__global__ void k_sp_broadcast(float* values, const unsigned* __restrict__ par, size_t n_vox) {
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g >= n_vox) return;
    const unsigned r = par[g];
    if (r == SP_EMPTY || r == (unsigned)g) return;
    values[g] = copysignf(values[g], values[r]);
}

}  // namespace

bs_status bs_mesh_closed_impl(bs_context* ctx, const float* d_tris, size_t n_tris, bool* closed) {
    cudaStream_t st = ctx->stream;
    *closed = false;
    const size_t nv = n_tris * 3;
    if (n_tris == 0 || nv >= 0x7FFFFFFFull) return BS_OK;
    size_t cap = 1024; while (cap < 2 * nv) cap <<= 1;
    unsigned *table = nullptr, *vid = nullptr; u64 *k1 = nullptr, *k2 = nullptr; int* d_bad = nullptr; void* d_tmp = nullptr; size_t tmp = 0;
    BS_TRY(bs_alloc(ctx, &table, cap)); BS_TRY(bs_alloc(ctx, &vid, nv)); BS_TRY(bs_alloc(ctx, &k1, nv)); BS_TRY(bs_alloc(ctx, &k2, nv)); BS_TRY(bs_alloc(ctx, &d_bad, 1));
    BS_CUDA(ctx, cudaMemsetAsync(table, 0xFF, cap * sizeof(unsigned), st));
    BS_CUDA(ctx, cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    bs_count_launch(), k_sp_vertex_ids<<<bs_blocks(nv, 256), 256, 0, st>>>(d_tris, nv, table, (unsigned)(cap - 1), vid);
    bs_count_launch(), k_sp_edges<<<bs_blocks(n_tris, 256), 256, 0, st>>>(vid, n_tris, k1, d_bad);
    cub::DeviceRadixSort::SortKeys(nullptr, tmp, k1, k2, (int)nv, 0, 64, st);
    BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp));
    cub::DeviceRadixSort::SortKeys(d_tmp, tmp, k1, k2, (int)nv, 0, 64, st);
    bs_count_launch(), k_sp_edges_check<<<bs_blocks(nv, 256), 256, 0, st>>>(k2, nv, d_bad);
    int bad = 1;
    BS_CUDA(ctx, cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    BS_CUDA(ctx, cudaStreamSynchronize(st));
    bs_free(ctx, d_tmp); bs_free(ctx, table); bs_free(ctx, vid); bs_free(ctx, k1); bs_free(ctx, k2); bs_free(ctx, d_bad);
    *closed = bad == 0;
    return BS_OK;
}

// components of the band voxels of `vol` (values = unsigned distances, masks = active bits): *d_par[g] = representative
// voxel of g = brick * 512 + offset (0xFFFFFFFF for inactive voxels), *d_seed = masks of the representatives
bs_status bs_sign_components_impl(bs_context* ctx, const bs_volume* vol, unsigned** d_par, unsigned long long** d_seed) {
    cudaStream_t st = ctx->stream;
    const size_t n = vol->n_bricks;
    *d_par = nullptr; *d_seed = nullptr;
    if (n == 0 || n * 512 >= 0xFFFFFFFFull) return bs_fail(ctx, BS_ERR_RANGE, "sign propagation: too many voxels for 32-bit component ids");
    unsigned* par = nullptr; u64* seed = nullptr;
    BS_TRY(bs_alloc(ctx, &par, n * 512)); BS_TRY(bs_alloc(ctx, &seed, n * 8));
    bs_count_launch(), k_sp_bricks<<<(unsigned)n, 512, 0, st>>>(vol->values, vol->masks, vol->voxel_size, par);
    bs_count_launch(), k_sp_faces<<<(unsigned)n, 192, 0, st>>>(vol->keys, n, vol->values, vol->masks, vol->voxel_size, par);
    bs_count_launch(), k_sp_flatten<<<(unsigned)n, 512, 0, st>>>(par, seed);
    BS_CUDA(ctx, cudaGetLastError());
    *d_par = par; *d_seed = seed;
    return BS_OK;
}
void bs_sign_chunks_from_masks(bs_context* ctx, const unsigned long long* d_masks, size_t n_bricks, unsigned* d_nchunks, int per_chunk) {
    bs_count_launch(), k_sp_chunks<<<bs_blocks(n_bricks, 256), 256, 0, ctx->stream>>>(d_masks, n_bricks, d_nchunks, per_chunk);
}
bs_status bs_sign_broadcast_impl(bs_context* ctx, bs_volume* vol, const unsigned* d_par) {
    const size_t nv = vol->n_bricks * 512;
    if (nv) bs_count_launch(), k_sp_broadcast<<<bs_blocks(nv, 256), 256, 0, ctx->stream>>>(vol->values, d_par, nv);
    BS_CUDA(ctx, cudaGetLastError());
    return BS_OK;
}
