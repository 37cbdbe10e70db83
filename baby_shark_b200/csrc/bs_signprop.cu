// Sign propagation for closed meshes (sm_100a): the default sign path of MeshToVolume::compute_sings
// (src/voxel/mesh_to_volume.rs:198-281) whenever the input is a closed 2-cycle; open meshes keep the per-voxel
// winding-number traversal of bs_fwn.cu, which is the reference's semantics there.
//
// The reference evaluates the (approximate) winding number of EVERY active voxel and thresholds it at 0.2. On a closed
// mesh -- every directed edge a->b is matched by an edge b->a, i.e. the boundary of the triangle chain is zero -- the
// exact winding number is an integer that only changes across the surface, so it is the same for two lattice neighbours
// p, q whenever the segment pq cannot meet the surface. That is certified from the unsigned distances already computed:
//     min(d_p, vs) + min(d_q, vs) > |pq| + tol
// (the scatter-min distance of bs_convert.cu is exact below one voxel: a sub-triangle whose integer box misses a lattice
// point is at least one voxel away from it). `tol` (bs_context::sp_tol, computed per mesh by k_tri_counts) bounds what
// f32 rounding can move: the sub-triangles the distances were measured to come from running sums of up to n+1 f32
// additions (mesh_to_volume.rs:90-115) and drift from the true triangle -- which is what the winding number sees -- by at
// most 1.74 (n + 13) 2^-24 max|coord|, and the lattice positions / box roundings add a few ulps of the coordinates.
//
// Steps: (1) closed? (2) union-find over the certified face links of the band voxels -- inside a brick in shared memory,
// across brick faces in global memory with path halving; (3) the per-voxel traversal of bs_fwn.cu runs on ONE voxel per
// component (two shell components plus the 5-10 % of voxels that hug the surface and certify no link); (4) every other
// voxel copies the sign of its component's representative.
//
// (1) has two implementations. Default: a 128-bit multiset fingerprint -- sum over directed edges of H(a, b) must equal
// the sum of H(b, a), H a 2 x 64-bit mix of the six coordinate words (-0 folded into +0; any non-finite coordinate means
// "not closed"): one streaming pass over the triangles (0.1 ms for 10 M) instead of a hash table + 64-bit sort of 30 M
// edge keys (3.6 ms); a false "closed" needs a 128-bit collision. BSHARK_CLOSED_CHECK=exact selects the sort-based test
// (vertices identified by exact coordinates like merge_points, every directed edge exactly once and its reverse exactly
// once); tests/test_gpu_signprop.py runs both on every test mesh.
#include "bs_common.cuh"
#include <cub/cub.cuh>
#include <cstdlib>
#include <cstring>

namespace {

typedef unsigned long long u64;
constexpr unsigned SP_EMPTY = 0xFFFFFFFFu;

// ---- (1a) closed? multiset fingerprint --------------------------------------------------------------------------------
__device__ __forceinline__ u64 sp_mix(u64 h) { h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33; return h; }
__device__ __forceinline__ void sp_edge_hash(const unsigned* a, const unsigned* b, u64& h0, u64& h1) {
    const u64 w0 = (u64)a[0] | ((u64)a[1] << 32), w1 = (u64)a[2] | ((u64)b[0] << 32), w2 = (u64)b[1] | ((u64)b[2] << 32);
    u64 x = sp_mix(w0 ^ 0x9E3779B97F4A7C15ULL); x = sp_mix(x ^ w1); x = sp_mix(x + w2);
    u64 y = sp_mix(w2 ^ 0xD6E8FEB86659FD93ULL); y = sp_mix(y + w0); y = sp_mix(y ^ (w1 * 0x9FB21C651E98DF25ULL));
    h0 = x; h1 = y;
}
__global__ void __launch_bounds__(256) k_sp_fingerprint(const float* __restrict__ tris, size_t n_tris, u64* out /*[4] fwd0 fwd1 rev0 rev1*/, int* bad) {
    u64 f0 = 0, f1 = 0, r0 = 0, r1 = 0; bool nonfinite = false;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n_tris; t += (size_t)gridDim.x * blockDim.x) {
        unsigned v[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const float f = tris[9 * t + i] + 0.0f;  // -0 -> +0
            if (!(fabsf(f) <= 3.0e38f)) nonfinite = true;
            v[i] = __float_as_uint(f);
        }
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            const unsigned* a = v + 3 * e; const unsigned* b = v + 3 * ((e + 1) % 3);
            u64 h0, h1;
            sp_edge_hash(a, b, h0, h1); f0 += h0; f1 += h1;
            sp_edge_hash(b, a, h0, h1); r0 += h0; r1 += h1;
        }
    }
    typedef cub::BlockReduce<u64, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    u64 s;
    s = BR(tmp).Sum(f0); if (threadIdx.x == 0) atomicAdd(out + 0, s); __syncthreads();
    s = BR(tmp).Sum(f1); if (threadIdx.x == 0) atomicAdd(out + 1, s); __syncthreads();
    s = BR(tmp).Sum(r0); if (threadIdx.x == 0) atomicAdd(out + 2, s); __syncthreads();
    s = BR(tmp).Sum(r1); if (threadIdx.x == 0) atomicAdd(out + 3, s);
    if (nonfinite) *bad = 1;
}

// ---- (1b) closed? exact test ---------------------------------------------------------------------------------------------
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
// shared-memory phase (one brick): plain find, the trees are shallow
__device__ __forceinline__ unsigned sp_find_s(volatile unsigned* par, unsigned x) {
    unsigned p;
    while ((p = par[x]) != x) x = p;
    return x;
}
__device__ __forceinline__ void sp_union_s(unsigned* par, unsigned a, unsigned b) {
    for (;;) {
        a = sp_find_s(par, a); b = sp_find_s(par, b);
        if (a == b) return;
        if (a > b) { const unsigned t = a; a = b; b = t; }
        if (atomicCAS(par + b, b, a) == b) return;
    }
}
// global phase: the two shell components span every brick, so finds halve the path as they go (a stale or concurrent
// write only ever replaces a parent by one of its ancestors; roots change by atomicCAS alone). Loads bypass L1.
__device__ __forceinline__ unsigned sp_find_g(unsigned* par, unsigned x) {
    volatile unsigned* vp = par;
    unsigned p = vp[x];
    while (p != x) {
        const unsigned g = vp[p];
        if (g != p) vp[x] = g;
        x = p; p = g;
    }
    return x;
}
__device__ __forceinline__ void sp_union_g(unsigned* par, unsigned a, unsigned b) {
    for (;;) {
        a = sp_find_g(par, a); b = sp_find_g(par, b);
        if (a == b) return;
        if (a > b) { const unsigned t = a; a = b; b = t; }
        if (atomicCAS(par + b, b, a) == b) return;
    }
}
// one CTA (512 threads, one per voxel) per brick: components of the brick's own certified links, written as global ids
__global__ void __launch_bounds__(512) k_sp_bricks(const float* __restrict__ values, const u64* __restrict__ masks, float vs, float thr, unsigned* par) {
    __shared__ unsigned s_par[512];
    __shared__ float s_cap[512];
    const size_t b = blockIdx.x;
    const unsigned t = threadIdx.x;
    const bool act = (masks[b * 8 + (t >> 6)] >> (t & 63)) & 1ull;
    const float cap = act ? fminf(fabsf(values[b * 512 + t]), vs) : -1.0f;
    s_par[t] = t; s_cap[t] = cap;
    __syncthreads();
    if (act) {
        const unsigned x = t >> 6, y = (t >> 3) & 7, z = t & 7;
        if (x < 7 && s_cap[t + 64] >= 0.f && cap + s_cap[t + 64] > thr) sp_union_s(s_par, t, t + 64);
        if (y < 7 && s_cap[t + 8] >= 0.f && cap + s_cap[t + 8] > thr) sp_union_s(s_par, t, t + 8);
        if (z < 7 && s_cap[t + 1] >= 0.f && cap + s_cap[t + 1] > thr) sp_union_s(s_par, t, t + 1);
    }
    __syncthreads();
    par[b * 512 + t] = act ? (unsigned)(b * 512) + sp_find_s(s_par, t) : SP_EMPTY;
}
__device__ __forceinline__ long long sp_find_key(const u64* keys, size_t n, u64 k) {
    size_t lo = 0, hi = n;
    while (lo < hi) { const size_t mid = (lo + hi) >> 1; if (keys[mid] < k) lo = mid + 1; else hi = mid; }
    return (lo < n && keys[lo] == k) ? (long long)lo : -1;
}
// one CTA (192 threads) per brick: certified links across its +x, +y, +z faces. Most of a face's links join the same
// two brick-level roots: a lane skips its union when the previous lane of its warp holds the same pair.
__global__ void __launch_bounds__(192) k_sp_faces(const u64* __restrict__ keys, size_t n, const float* __restrict__ values, const u64* __restrict__ masks, float vs, float thr, unsigned* par) {
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
    const long long nb = s_nb[ax];  // uniform per warp (64 threads per axis)
    if (nb < 0) return;
    const unsigned op = ax == 0 ? ((7u << 6) | (u << 3) | v) : (ax == 1 ? ((u << 6) | (7u << 3) | v) : ((u << 6) | (v << 3) | 7u));
    const unsigned oq = ax == 0 ? ((u << 3) | v) : (ax == 1 ? ((u << 6) | v) : ((u << 6) | (v << 3)));
    const bool ap = (masks[b * 8 + (op >> 6)] >> (op & 63)) & 1ull, aq = (masks[(size_t)nb * 8 + (oq >> 6)] >> (oq & 63)) & 1ull;
    bool link = false; unsigned gp = 0, gq = 0, rp = SP_EMPTY, rq = SP_EMPTY;
    if (ap && aq) {
        const float cp = fminf(fabsf(values[b * 512 + op]), vs), cq = fminf(fabsf(values[(size_t)nb * 512 + oq]), vs);
        link = cp + cq > thr;
        gp = (unsigned)(b * 512 + op); gq = (unsigned)((size_t)nb * 512 + oq);
        if (link) { rp = __ldcg(par + gp); rq = __ldcg(par + gq); }  // brick-level roots (or already something above them)
    }
    const unsigned lane = threadIdx.x & 31;
    const unsigned prp = __shfl_up_sync(0xFFFFFFFFu, rp, 1), prq = __shfl_up_sync(0xFFFFFFFFu, rq, 1);
    if (link && !(lane > 0 && prp == rp && prq == rq)) sp_union_g(par, gp, gq);
}
// flatten + seed masks: a voxel is its component's representative iff it is its own root
__global__ void __launch_bounds__(512) k_sp_flatten(unsigned* par, u64* seed_masks, unsigned* n_chunks, int per_chunk, u64* n_seeds) {
    __shared__ unsigned s_cnt;
    const size_t b = blockIdx.x;
    const unsigned t = threadIdx.x;
    if (t == 0) s_cnt = 0;
    __syncthreads();
    const unsigned g = (unsigned)(b * 512 + t);
    bool seed = false;
    // read-only find: a path-halving write by another thread could land AFTER this thread's `par[g] = r` and leave g
    // pointing at an intermediate ancestor (seen at 2048^3: 31 voxels of 36 M took the sign of a non-representative)
    if (par[g] != SP_EMPTY) { unsigned r = g, p; volatile unsigned* vp = par; while ((p = vp[r]) != r) r = p; par[g] = r; seed = r == g; }
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, seed);
    if ((t & 31) == 0) { reinterpret_cast<unsigned*>(seed_masks + b * 8)[t >> 5] = bal; if (bal) atomicAdd(&s_cnt, (unsigned)__popc(bal)); }  // bit t of the brick's 512-bit mask
    __syncthreads();
    if (t == 0) { n_chunks[b] = (s_cnt + per_chunk - 1) / per_chunk; if (s_cnt) atomicAdd(n_seeds, (u64)s_cnt); }
}
// (4) every non-representative voxel takes the sign its representative got from the traversal
__global__ void k_sp_broadcast(float* values, const unsigned* __restrict__ par, size_t n_vox) {
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g >= n_vox) return;
    const unsigned r = par[g];
    if (r == SP_EMPTY || r == (unsigned)g) return;
    values[g] = copysignf(values[g], __ldcg(values + r));
}

}  // namespace

static bs_status mesh_closed_exact(bs_context* ctx, const float* d_tris, size_t n_tris, bool* closed) {
    cudaStream_t st = ctx->stream;
    const size_t nv = n_tris * 3;
    if (nv >= 0x7FFFFFFFull) return BS_OK;
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

// Starts the closedness test on the stream; bs_mesh_closed_finish reads the verdict (after the caller's next sync)
bs_status bs_mesh_closed_begin(bs_context* ctx, const float* d_tris, size_t n_tris, bs_closed_check* chk) {
    cudaStream_t st = ctx->stream;
    chk->exact = false; chk->closed = false; chk->d_sums = nullptr; chk->d_bad = nullptr; chk->pending = false;
    if (n_tris == 0) return BS_OK;
    const char* e = getenv("BSHARK_CLOSED_CHECK");
    if (e && strcmp(e, "exact") == 0) { chk->exact = true; return mesh_closed_exact(ctx, d_tris, n_tris, &chk->closed); }
    BS_TRY(bs_alloc(ctx, &chk->d_sums, 4)); BS_TRY(bs_alloc(ctx, &chk->d_bad, 1));
    BS_CUDA(ctx, cudaMemsetAsync(chk->d_sums, 0, 4 * sizeof(u64), st));
    BS_CUDA(ctx, cudaMemsetAsync(chk->d_bad, 0, sizeof(int), st));
    const unsigned grid = (unsigned)std::min<size_t>(bs_blocks(n_tris, 256), (size_t)ctx->sm_count * 16);
    bs_count_launch(), k_sp_fingerprint<<<grid, 256, 0, st>>>(d_tris, n_tris, chk->d_sums, chk->d_bad);
    BS_CUDA(ctx, cudaMemcpyAsync(chk->h_sums, chk->d_sums, 4 * sizeof(u64), cudaMemcpyDeviceToHost, st));
    BS_CUDA(ctx, cudaMemcpyAsync(&chk->h_bad, chk->d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    chk->pending = true;
    return BS_OK;
}
// call after a stream synchronisation that follows bs_mesh_closed_begin
bool bs_mesh_closed_finish(bs_context* ctx, bs_closed_check* chk) {
    if (chk->pending) {
        chk->closed = chk->h_bad == 0 && chk->h_sums[0] == chk->h_sums[2] && chk->h_sums[1] == chk->h_sums[3];
        bs_free(ctx, chk->d_sums); bs_free(ctx, chk->d_bad);
        chk->d_sums = nullptr; chk->d_bad = nullptr; chk->pending = false;
    }
    return chk->closed;
}

// components of the band voxels of `vol` (values = unsigned distances, masks = active bits): *d_par[g] = representative
// voxel of g = brick * 512 + offset (0xFFFFFFFF for inactive voxels), *d_seed = masks of the representatives,
// d_nchunks[b] = work items (of per_chunk representatives) of brick b. *applicable = false (nothing allocated) when the
// volume is too large for 32-bit voxel ids or the rounding tolerance leaves no certifiable link: the caller keeps the
// per-voxel path.
bs_status bs_sign_components_impl(bs_context* ctx, const bs_volume* vol, float tol, unsigned** d_par, unsigned long long** d_seed, unsigned* d_nchunks, int per_chunk, unsigned long long* d_nseeds, bool* applicable) {
    cudaStream_t st = ctx->stream;
    const size_t n = vol->n_bricks;
    *d_par = nullptr; *d_seed = nullptr; *applicable = false;
    const float vs = vol->voxel_size;
    if (n == 0 || n * 512 >= 0xFFFFFFFFull || !(tol >= 0.f) || !(tol < 0.5f * vs)) return BS_OK;
    const float thr = vs + tol;
    unsigned* par = nullptr; u64* seed = nullptr;
    BS_TRY(bs_alloc(ctx, &par, n * 512)); BS_TRY(bs_alloc(ctx, &seed, n * 8));
    bs_count_launch(), k_sp_bricks<<<(unsigned)n, 512, 0, st>>>(vol->values, vol->masks, vs, thr, par);
    bs_count_launch(), k_sp_faces<<<(unsigned)n, 192, 0, st>>>(vol->keys, n, vol->values, vol->masks, vs, thr, par);
    bs_count_launch(), k_sp_flatten<<<(unsigned)n, 512, 0, st>>>(par, seed, d_nchunks, per_chunk, d_nseeds);
    BS_CUDA(ctx, cudaGetLastError());
    *d_par = par; *d_seed = seed; *applicable = true;
    return BS_OK;
}
bs_status bs_sign_broadcast_impl(bs_context* ctx, bs_volume* vol, const unsigned* d_par) {
    const size_t nv = vol->n_bricks * 512;
    if (nv) bs_count_launch(), k_sp_broadcast<<<bs_blocks(nv, 256), 256, 0, ctx->stream>>>(vol->values, d_par, nv);
    BS_CUDA(ctx, cudaGetLastError());
    return BS_OK;
}
