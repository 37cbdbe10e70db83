// Inside/outside signs by fast winding numbers over a device-built LBVH (sm_100a).
// Replaces WindingNumbers::{from_mesh, approximate} (src/spatial_partitioning/aabb_tree.rs:636-691,711-816)
// and MeshToVolume::compute_sings / ComputeSignsVisitor (src/voxel/mesh_to_volume.rs:198-281):
//     value = copysign(|d|, wn(p) < 0.2 ? +1 : -1),   p = idx as f32 * voxel_size.
//
// The reference builds a serial top-down binned-SAH tree with <= 3 triangles per leaf. Here: 48-bit Morton codes
// of the triangle centroids, radix sort, leaves of LEAF (= 1) consecutive triangles, Karras' (2012) binary radix tree
// emitted fully in parallel, and one bottom-up pass (second-arrival atomics, in shared memory for the 94 % of the
// parents whose leaves belong to one CTA) for the per-node moments the reference keeps (aabb_tree.rs:723-801): area-weighted normal (order 1), sum(area c n^T) - p~ (sum area n)^T
// (order 2), dipole centre p~ = sum(area c)/sum(area). The node radius is the distance from p~ to the farthest
// corner of the node's box (the reference takes the farther of the two extreme corners only, :734-736; this one
// is a true bound, so the far-field test is never looser than the reference's).
//
// Traversal is warp-cooperative: a warp owns 32 Morton-adjacent active voxels of ONE brick and walks ONE shared
// stack of (node, lane mask) entries (see WarpWinding): per-voxel acceptance exactly as the reference's
// criterion (|p - p~| > 2 radius, :666), warp-uniform control flow, broadcast loads of 256-byte traversal records
// (a node's four grandchildren, interleaved in pairs for packed f32x2 evaluation). The top of the tree is walked once
// per brick (k_brick_pass): nodes far from the whole brick are sampled at 27 points and interpolated per voxel.
// Sharded runs split the work items of bricks under dense slivers by triangles (k_expand_roots, k_sign_finish).
// Only the 0.2 threshold matters downstream, so FMA contraction is allowed here (unlike the distance stage).
#include "bs_common.cuh"
#include <cub/cub.cuh>
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace {

#ifndef BS_LEAF
#define BS_LEAF 1
#endif
#ifndef BS_VPL
#define BS_VPL 1
#endif
// Traversal records keep their entries interleaved in pairs and k_sign evaluates a pair with packed f32x2 instructions
// (FFMA2 / FMUL2 / FADD2, new on sm_100): 36.4 -> 33.3 ms on config 5. -DBS_NO_PAIRS restores the scalar path.
#if BS_VPL == 1 && !defined(BS_NO_PAIRS) && !defined(BS_PAIRS)
#define BS_PAIRS
#endif
constexpr int LEAF = BS_LEAF;     // triangles per leaf
constexpr int STACK = 208;        // per-warp stack entries: sub-tree roots (<= MAX_ROOTS = 96) + tree depth (<= 63 + 32 with index tie-breaks)
constexpr int WARPS_PER_BLOCK = 4;
constexpr float BETA = 2.0f;      // accuracy_scale (mesh_to_volume.rs:264)
constexpr float INV_4PI = 0.07957747154594767f;

struct Raw {  // additive moments of a node (aabb_tree.rs:694-700) + box
    float area, awc[3], awn[3], o1sum[9], bbmin[3], bbmax[3], pad[2];
};
static_assert(sizeof(Raw) == 96, "Raw is 24 floats");

struct Tree {
    // node ids: internal [0, n-1), leaves [n-1, 2n-1)
    const float4* hdr;      // per node {p~.xyz, beta^2 r^2}
    const float4* coef;     // per node 3 x float4 far-field coefficients
    const float4* rec;      // per internal node: traversal record (k_records)
    const float4* tris;     // 3 x float4 per sorted triangle (9 floats + pad), LEAF per leaf, padded with degenerate triangles
    unsigned n_leaves;
    unsigned root;
    unsigned* derr;         // bs_context::d_err
};

__device__ __forceinline__ int f2ord(float f) { int b = __float_as_int(f); return b >= 0 ? b : b ^ 0x7FFFFFFF; }
__device__ __forceinline__ float ord2f(int o) { return __int_as_float(o >= 0 ? o : o ^ 0x7FFFFFFF); }

__global__ void k_centroid_bounds(const float* __restrict__ tris, size_t n, int* bounds /*min xyz, max xyz as ordered ints*/) {
    // grid-stride, then warp shuffle + shared-memory reduction: six atomics per CTA, not per warp
    __shared__ int s_lo[3][8], s_hi[3][8];
    int lo[3] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF}, hi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        const float* p = tris + 9 * t;
        float c[3];
        for (int d = 0; d < 3; ++d) c[d] = (p[d] + p[3 + d] + p[6 + d]) * (1.0f / 3.0f);
        const bool ok = (c[0] == c[0]) && (c[1] == c[1]) && (c[2] == c[2]) && fabsf(c[0]) < 1e30f && fabsf(c[1]) < 1e30f && fabsf(c[2]) < 1e30f;
        if (ok) for (int d = 0; d < 3; ++d) { const int o = f2ord(c[d]); lo[d] = min(lo[d], o); hi[d] = max(hi[d], o); }
    }
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int d = 0; d < 3; ++d) {
        for (int o = 16; o; o >>= 1) { lo[d] = min(lo[d], __shfl_xor_sync(0xFFFFFFFFu, lo[d], o)); hi[d] = max(hi[d], __shfl_xor_sync(0xFFFFFFFFu, hi[d], o)); }
        if (lane == 0) { s_lo[d][w] = lo[d]; s_hi[d][w] = hi[d]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int d = threadIdx.x;
        int l = s_lo[d][0], h = s_hi[d][0];
        for (unsigned k = 1; k < blockDim.x / 32; ++k) { l = min(l, s_lo[d][k]); h = max(h, s_hi[d][k]); }
        atomicMin(bounds + d, l); atomicMax(bounds + 3 + d, h);
    }
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long v) {
    v &= 0x1FFFFFull;
    v = (v | v << 32) & 0x1F00000000FFFFull; v = (v | v << 16) & 0x1F0000FF0000FFull;
    v = (v | v << 8) & 0x100F00F00F00F00Full; v = (v | v << 4) & 0x10C30C30C30C30C3ull; v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}
__global__ void k_morton(const float* __restrict__ tris, size_t n, const int* bounds, unsigned long long* codes, unsigned* ids) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float* p = tris + 9 * t;
    unsigned long long code = 0;
    for (int d = 0; d < 3; ++d) {
        float lo = ord2f(bounds[d]), hi = ord2f(bounds[3 + d]);
        float c = (p[d] + p[3 + d] + p[6 + d]) * (1.0f / 3.0f);
        float ext = hi - lo;
        float u = ext > 0.f ? (c - lo) / ext : 0.f;
        if (!(u >= 0.f)) u = 0.f;
        if (u > 1.f) u = 1.f;
        unsigned long long q = (unsigned long long)(u * 65535.0f);  // 16 bits per axis: 32 cells per voxel at 2048^3, two radix passes fewer than 21
        code |= spread21(q) << (2 - d);
    }
    codes[t] = code; ids[t] = (unsigned)t;
}

// ---- Karras 2012: one thread per internal node -----------------------------------------------------------------
// key of leaf g = Morton code of its first triangle; ties broken by the leaf index.
__device__ __forceinline__ int delta(const unsigned long long* __restrict__ codes, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = codes[(size_t)i * LEAF], b = codes[(size_t)j * LEAF];
    if (a != b) return __clzll((long long)(a ^ b));
    return 64 + __clz(i ^ j);
}
__global__ void k_karras(const unsigned long long* __restrict__ codes, int n, int* left, int* right, int* parent, int* other /*node i covers leaves [min(i, other), max(i, other)]*/) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(codes, n, i, i + 1) - delta(codes, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(codes, n, i, i - d);
    int lmax = 2;
    while (delta(codes, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1) if (delta(codes, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(codes, n, i, j);
    int s = 0, t = l;
    do { t = (t + 1) >> 1; if (delta(codes, n, i, i + (s + t) * d) > dnode) s += t; } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int lc = (lo == gamma) ? (n - 1 + gamma) : gamma;
    const int rc = (hi == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
    left[i] = lc; right[i] = rc; other[i] = j;
    parent[lc] = i; parent[rc] = i;
    if (i == 0) parent[0] = -1;
}

__device__ __forceinline__ void raw_load_cg(const Raw* p, Raw& r) {
    const float4* s = reinterpret_cast<const float4*>(p); float4* d = reinterpret_cast<float4*>(&r);
#pragma unroll
    for (int i = 0; i < 6; ++i) d[i] = __ldcg(s + i);
}

// Per node: header {p~.xyz, beta^2 r^2} and far-field coefficients {o1.xyz, trM} {m00 m11 m22 m01+m10} {m02+m20 m12+m21 - -}
__device__ __forceinline__ void finalize_node(const Raw& r, size_t i, float4* hdr, float4* coef) {
    float c[3], rad2 = 0.f;
    for (int d = 0; d < 3; ++d) {
        c[d] = r.awc[d] / r.area;  // NaN for an all-degenerate node: never "far", always descended (as in the reference)
        const float e = fmaxf(fabsf(r.bbmin[d] - c[d]), fabsf(r.bbmax[d] - c[d]));
        rad2 += e * e;
    }
    if (r.pad[0] > 0.f && r.pad[0] * r.pad[0] < rad2) rad2 = r.pad[0] * r.pad[0];  // the tighter of two valid bounds
    float m[9];  // m[col*3+row]
    for (int col = 0; col < 3; ++col) for (int row = 0; row < 3; ++row) m[col * 3 + row] = r.o1sum[col * 3 + row] - c[row] * r.awn[col];
    hdr[i] = make_float4(c[0], c[1], c[2], BETA * BETA * rad2);
    coef[3 * (size_t)i + 0] = make_float4(r.awn[0], r.awn[1], r.awn[2], m[0] + m[4] + m[8]);
    coef[3 * (size_t)i + 1] = make_float4(m[0], m[4], m[8], m[1] + m[3]);
    coef[3 * (size_t)i + 2] = make_float4(m[2] + m[6], m[5] + m[7], 0.f, 0.f);
}

// moments of a node from its two children (aabb_tree.rs:779-801); a = left (accumulates), b = right
__device__ __forceinline__ void merge_raw(Raw& a, const Raw& b) {
    a.area += b.area;
    for (int d = 0; d < 3; ++d) { a.awc[d] += b.awc[d]; a.awn[d] += b.awn[d]; a.bbmin[d] = fminf(a.bbmin[d], b.bbmin[d]); a.bbmax[d] = fmaxf(a.bbmax[d], b.bbmax[d]); }
    for (int j = 0; j < 9; ++j) a.o1sum[j] += b.o1sum[j];
    // |p~_child - p~| + r_child bounds the child's triangles about the merged centre
    const float ia = 1.0f / (a.area - b.area), ib = 1.0f / b.area, in = 1.0f / a.area;  // a.area is already the sum
    float da = 0.f, db = 0.f;
    for (int d = 0; d < 3; ++d) {
        const float cn = a.awc[d] * in, ca = (a.awc[d] - b.awc[d]) * ia, cb = b.awc[d] * ib;
        da += (ca - cn) * (ca - cn); db += (cb - cn) * (cb - cn);
    }
    const float ra = sqrtf(da) + a.pad[0], rb = sqrtf(db) + b.pad[0];
    a.pad[0] = (ra == ra && rb == rb) ? fmaxf(ra, rb) : 3.0e38f;  // a degenerate child: keep only the box bound
}

constexpr int CLIMB_TPB = 128;
// leaves: gather LEAF sorted triangles, store them for traversal, accumulate the leaf's moments
// (aabb_tree.rs:749-777), then climb: the second child to arrive at a parent combines both (:779-801).
__global__ void __launch_bounds__(CLIMB_TPB) k_leaves_and_climb(const float* __restrict__ tris, const unsigned* __restrict__ ids, size_t n_tris, float4* sorted, Raw* raw,
                                   const int* __restrict__ left, const int* __restrict__ right, const int* __restrict__ parent, const int* __restrict__ other,
                                   unsigned* flags, int n, float4* hdr, float4* coef) {
    __shared__ __align__(16) Raw s_raw[2 * CLIMB_TPB];  // [0, 128): the CTA's leaves, [128, 256): internal node i at 128 + i - first
    __shared__ unsigned s_flag[CLIMB_TPB];
    const int first = blockIdx.x * CLIMB_TPB;
    s_flag[threadIdx.x] = 0;
    __syncthreads();
    const int g = first + (int)threadIdx.x;
    if (g >= n) return;  // no block-wide barrier below
    Raw r;
    r.area = 0.f;
    for (int d = 0; d < 3; ++d) { r.awc[d] = 0.f; r.awn[d] = 0.f; r.bbmin[d] = 3.0e38f; r.bbmax[d] = -3.0e38f; }
    for (int i = 0; i < 9; ++i) r.o1sum[i] = 0.f;
    r.pad[0] = r.pad[1] = 0.f;
    for (int k = 0; k < LEAF; ++k) {
        const size_t s = (size_t)g * LEAF + k;
        float v[9];
        if (s < n_tris) { const float* p = tris + 9 * (size_t)ids[s]; for (int i = 0; i < 9; ++i) v[i] = p[i]; }
        else { for (int i = 0; i < 9; ++i) v[i] = 0.f; }  // padding: degenerate triangle, contributes nothing
        sorted[3 * s + 0] = make_float4(v[0], v[1], v[2], v[3]);
        sorted[3 * s + 1] = make_float4(v[4], v[5], v[6], v[7]);
        sorted[3 * s + 2] = make_float4(v[8], 0.f, 0.f, 0.f);
        if (s >= n_tris) continue;
        for (int d = 0; d < 3; ++d) {
            r.bbmin[d] = fminf(r.bbmin[d], fminf(v[d], fminf(v[3 + d], v[6 + d])));
            r.bbmax[d] = fmaxf(r.bbmax[d], fmaxf(v[d], fmaxf(v[3 + d], v[6 + d])));
        }
        const float e1[3] = {v[3] - v[0], v[4] - v[1], v[5] - v[2]}, e2[3] = {v[6] - v[0], v[7] - v[1], v[8] - v[2]};
        const float cr[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        const float n2 = cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2];
        if (!(n2 > 0.f)) continue;  // degenerate triangles are skipped (:757-759)
        const float len = sqrtf(n2), area = 0.5f * len;
        const float nn[3] = {cr[0] / len, cr[1] / len, cr[2] / len};
        const float c[3] = {(v[0] + v[3] + v[6]) / 3.0f, (v[1] + v[4] + v[7]) / 3.0f, (v[2] + v[5] + v[8]) / 3.0f};
        r.area += area;
        for (int d = 0; d < 3; ++d) { r.awn[d] += area * nn[d]; r.awc[d] += area * c[d]; }
        for (int col = 0; col < 3; ++col) for (int row = 0; row < 3; ++row) r.o1sum[col * 3 + row] += area * c[row] * nn[col];
    }
    {   // bounding radius about the dipole centre: exact over the leaf's vertices
        float c[3] = {r.awc[0] / r.area, r.awc[1] / r.area, r.awc[2] / r.area}, m2 = 0.f;
        for (int k = 0; k < LEAF; ++k) {
            const size_t s = (size_t)g * LEAF + k;
            if (s >= n_tris) break;
            const float* p = tris + 9 * (size_t)ids[s];
            for (int v = 0; v < 3; ++v) { const float dx = p[3 * v] - c[0], dy = p[3 * v + 1] - c[1], dz = p[3 * v + 2] - c[2]; m2 = fmaxf(m2, dx * dx + dy * dy + dz * dz); }
        }
        r.pad[0] = sqrtf(m2);  // NaN centre (all-degenerate leaf) -> fmaxf drops the NaNs -> 0, finalize falls back to the box bound
    }
    finalize_node(r, (size_t)(n - 1 + g), hdr, coef);  // header + far-field coefficients as soon as the moments are complete
    if (n == 1) { raw[0] = r; return; }
    // Climb. A parent whose whole leaf range lies inside this CTA's 128 leaves is handled in shared memory (block-scope
    // fences, shared atomics, no global traffic for the moments): ~94 % of the internal nodes. The first node whose
    // parent reaches outside the CTA switches to the global protocol for the rest of its path.
    int node = parent[n - 1 + g];
    int cur_id = n - 1 + g, cur_slot = (int)threadIdx.x;
    Raw cur = r;
    for (;;) {
        const int oj = other[node];
        const int lo = min(node, oj), hi = max(node, oj);
        if (!(lo >= first && hi < first + CLIMB_TPB)) break;
        s_raw[cur_slot] = cur;
        __threadfence_block();
        if (atomicAdd(&s_flag[node - first], 1u) == 0u) return;  // first arrival: the sibling subtree is not finished yet
        __threadfence_block();
        const int lc = left[node], rc = right[node];
        Raw a = s_raw[lc >= n - 1 ? lc - (n - 1) - first : CLIMB_TPB + lc - first];
        const Raw b = s_raw[rc >= n - 1 ? rc - (n - 1) - first : CLIMB_TPB + rc - first];
        merge_raw(a, b);
        finalize_node(a, (size_t)node, hdr, coef);
        cur = a; cur_id = node; cur_slot = CLIMB_TPB + node - first;
        node = parent[node];
        if (node < 0) return;  // the root fits in one CTA
    }
    raw[cur_id] = cur;
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(&flags[node], 1u) == 0u) return;
        __threadfence();
        Raw a, b;
        raw_load_cg(raw + left[node], a); raw_load_cg(raw + right[node], b);
        merge_raw(a, b);
        raw[node] = a;
        finalize_node(a, (size_t)node, hdr, coef);
        node = parent[node];
    }
}

// Traversal record of internal node X: its (up to 4) grandchildren -- or a child itself when that child is a leaf --
// with everything a visit needs, so one pop costs ONE dependent memory round trip for four node tests.
// 256 B = two 128 B lines: [0..3] entry headers, [4] entry node ids (int4, -1 = none), then 10 coefficient floats
// per entry {o1.xyz, trM, m00, m11, m22, m01+m10, m02+m20, m12+m21} packed from float 20 on.
constexpr int REC = 16;  // float4 per record
__global__ void __launch_bounds__(256) k_records(const int* __restrict__ left, const int* __restrict__ right, const float4* __restrict__ hdr, const float4* __restrict__ coef, int n, float4* rec) {
    // four consecutive lanes build one record (one entry each) in shared memory; the CTA's 64 records (16 KB, contiguous
    // in global memory) then leave with coalesced 128-bit stores
    __shared__ float4 s_rec[64 * REC];
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const int x = (int)(g >> 2), e = (int)(g & 3);
    const bool valid = x < n - 1;
    int ids[4] = {-1, -1, -1, -1}, k = 0;
    if (valid) {
        const int ch[2] = {left[x], right[x]};
        for (int c = 0; c < 2; ++c) {
            if (ch[c] >= n - 1) ids[k++] = ch[c];
            else { ids[k++] = left[ch[c]]; ids[k++] = right[ch[c]]; }
        }
    }
    float4* r = s_rec + (threadIdx.x >> 2) * REC;
    const int id = ids[e];
    float4 h = make_float4(0.f, 0.f, 0.f, 0.f), c0 = h, c1 = h, c2 = h;
    if (id >= 0) { h = hdr[id]; c0 = coef[3 * (size_t)id]; c1 = coef[3 * (size_t)id + 1]; c2 = coef[3 * (size_t)id + 2]; }
#ifdef BS_PAIRS
    // entries are interleaved pairwise -- (e0, e1) and (e2, e3) -- so that every 64-bit word holds the same quantity of
    // two entries: the traversal evaluates a pair with packed f32x2 instructions (FFMA2 / FMUL2 / FADD2)
    {
        float* f = reinterpret_cast<float*>(r);
        const int p = e >> 1, half = e & 1;
        const float q[14] = {h.x, h.y, h.z, h.w, c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y};
#pragma unroll
        for (int j = 0; j < 14; ++j) f[(j < 4 ? p * 8 + j * 2 : 20 + p * 20 + (j - 4) * 2) + half] = q[j];
    }
#else
    r[e] = h;
    float2* o = reinterpret_cast<float2*>(reinterpret_cast<float*>(r) + 20 + 10 * e);  // 8 B aligned
    o[0] = make_float2(c0.x, c0.y); o[1] = make_float2(c0.z, c0.w); o[2] = make_float2(c1.x, c1.y); o[3] = make_float2(c1.z, c1.w); o[4] = make_float2(c2.x, c2.y);
#endif
    if (e == 0) r[4] = make_float4(__int_as_float(ids[0]), __int_as_float(ids[1]), __int_as_float(ids[2]), __int_as_float(ids[3]));
    if (e == 1) r[15] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const size_t first = (size_t)blockIdx.x * 64;                       // first record of this CTA
    const size_t n_here = first < (size_t)(n - 1) ? min((size_t)64, (size_t)(n - 1) - first) : 0;
    float4* dst = rec + first * REC;
    for (unsigned i = threadIdx.x; i < n_here * REC; i += 256) dst[i] = s_rec[i];
}

__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// atan2 to ~1e-5 rad (odd minimax polynomial on [0,1] + octant folding): the sum of solid angles only has to
// resolve the 0.2 threshold, libm's 1-ulp atan2f costs three times the instructions
__device__ __forceinline__ float fast_atan2(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float a = mn * fast_rcp(mx);
    const float s = a * a;
    float r = fmaf(fmaf(fmaf(fmaf(0.0208351f, s, -0.0851330f), s, 0.1801410f), s, -0.3302995f), s, 0.9998660f) * a;
    if (ay > ax) r = 1.57079637f - r;
    if (x < 0.f) r = 3.14159274f - r;
    return copysignf(r, y);
}

// solid_angle / (4 pi) (aabb_tree.rs:582-628), Van Oosterom-Strackee form
__device__ __forceinline__ float tri_winding(const float4 t0, const float4 t1, const float4 t2, float qx, float qy, float qz) {
    const float ax = t0.x - qx, ay = t0.y - qy, az = t0.z - qz;
    const float bx = t0.w - qx, by = t1.x - qy, bz = t1.y - qz;
    const float cx = t1.z - qx, cy = t1.w - qy, cz = t2.x - qz;
    const float la = fast_sqrt(fmaf(az, az, fmaf(ay, ay, ax * ax))), lb = fast_sqrt(fmaf(bz, bz, fmaf(by, by, bx * bx))), lc = fast_sqrt(fmaf(cz, cz, fmaf(cy, cy, cx * cx)));
    if (la == 0.f || lb == 0.f || lc == 0.f) return 0.f;
    const float det = ax * (by * cz - bz * cy) + ay * (bz * cx - bx * cz) + az * (bx * cy - by * cx);
    if (det == 0.f) return 0.f;
    const float den = la * lb * lc + (ax * bx + ay * by + az * bz) * lc + (ax * cx + ay * cy + az * cz) * lb + (bx * cx + by * cy + bz * cz) * la;
    return fast_atan2(det, den) * (2.0f * INV_4PI);
}

// Warp-cooperative winding numbers of up to 32*VPL query points (VPL per lane). The warp walks ONE shared stack
// of (node, lane masks) entries: a (lane, slot) takes part in a node only if none of its ancestors was already
// accepted as far for it, so every voxel gets exactly the sum its own traversal would give (criterion
// |p - p~| > 2 radius per voxel, aabb_tree.rs:666) while control flow stays warp-uniform and every node read is a
// warp-uniform (broadcast) load. The work issued is the union of the traversals, which for Morton-adjacent voxels
// of one brick is close to a single traversal.
// COUNT: also tally per-voxel node visits / far-field evaluations / exact triangle evaluations (cnt[0..2]).
template <bool COUNT, int VPL>
struct WarpWinding {
    const Tree& T;
    float qx[VPL], qy[VPL], qz[VPL], wn[VPL];
    unsigned* stack;   // STACK entries of (1 + VPL) words
    unsigned* cnt;
    int sp;
    unsigned lane;

    __device__ __forceinline__ void visit(unsigned id, const float4 h, const float4 c0, const float4 c1, const float4 c2, const unsigned* m) {
        unsigned near_m[VPL]; unsigned any_near = 0;
        if (COUNT && lane == 0) cnt[3]++;
        const unsigned lane_bit = 1u << lane;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const bool in = (m[v] & lane_bit) != 0;
            const float rx = h.x - qx[v], ry = h.y - qy[v], rz = h.z - qz[v];
            const float r2 = fmaf(rz, rz, fmaf(ry, ry, rx * rx));
            const bool far = in && (r2 > h.w);
            const unsigned far_m = __ballot_sync(0xFFFFFFFFu, far);
            near_m[v] = m[v] & ~far_m; any_near |= near_m[v];
            if (COUNT && in) cnt[0]++;
            if (far_m) {
                // far-field dipole (aabb_tree.rs:667-669, hessians :803-816):
                //   o1 . r/(4 pi |r|^3) + M : (I/(4 pi |r|^3) - 3 r r^T/(4 pi |r|^5)),  r = p~ - q
                float inv_r;
                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv_r) : "f"(r2));
                const float ir2 = inv_r * inv_r;
                const float k = (INV_4PI * inv_r) * ir2;
                const float d = fmaf(c0.x, rx, fmaf(c0.y, ry, fmaf(c0.z, rz, c0.w)));      // o1 . r + tr M
                const float t1 = fmaf(c1.w, ry, fmaf(c2.x, rz, c1.x * rx));                // m00 rx + (m01+m10) ry + (m02+m20) rz
                const float t2 = fmaf(c2.y, rz, c1.y * ry);                                // m11 ry + (m12+m21) rz
                const float rMr = fmaf(rx, t1, fmaf(ry, t2, (c1.z * rz) * rz));
                const float f = k * fmaf(-3.0f * ir2, rMr, d);
                if (far) { wn[v] += f; if (COUNT) cnt[1]++; }
            }
        }
        if (any_near == 0) return;
        // near for some voxel: internal nodes are expanded, leaves evaluated exactly -- both when popped, so the
        // (large) exact-evaluation code exists once instead of once per inlined visit
        if (sp >= STACK) { if (lane == 0) atomicOr(T.derr, BS_DERR_STACK); return; }  // reported as BS_ERR_RANGE by bs_convert_impl
        {
            if (id < T.n_leaves - 1) {  // start pulling the record this entry will need (two 128 B lines)
                const char* nxt = reinterpret_cast<const char*>(T.rec + (size_t)id * REC);
                asm volatile("prefetch.global.L1 [%0];" ::"l"(nxt));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(nxt + 128));
            }
#ifdef BS_UNIFORM_PUSH
            {   // every lane stores the same words to the same address: one wavefront, no divergent region
                unsigned* e = stack + sp * (1 + VPL);
                e[0] = id;
#pragma unroll
                for (int v = 0; v < VPL; ++v) e[1 + v] = near_m[v];
            }
#else
            if (lane == 0) {
                unsigned* e = stack + sp * (1 + VPL);
                e[0] = id;
#pragma unroll
                for (int v = 0; v < VPL; ++v) e[1 + v] = near_m[v];
            }
#endif
            ++sp;
        }
    }

    // leaf: sum of solid angles over its triangles (aabb_tree.rs:680-683) for the voxels in m
    __device__ __forceinline__ void exact_leaf(unsigned id, const unsigned* m) {
        const float4* t = T.tris + 3 * (size_t)(id - (T.n_leaves - 1)) * LEAF;
#pragma unroll
        for (int k = 0; k < LEAF; ++k) {
            const float4 t0 = __ldg(t + 3 * k), t1 = __ldg(t + 3 * k + 1), t2 = __ldg(t + 3 * k + 2);
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                if (m[v] == 0) continue;  // uniform
                const float e = tri_winding(t0, t1, t2, qx[v], qy[v], qz[v]);
                if ((m[v] >> lane) & 1) { wn[v] += e; if (COUNT) cnt[2]++; }
            }
        }
    }

#ifdef BS_PAIRS
    __device__ __forceinline__ void push(unsigned id, unsigned near_m) {
        if (near_m == 0) return;
        if (sp >= STACK) { if (lane == 0) atomicOr(T.derr, BS_DERR_STACK); return; }  // reported as BS_ERR_RANGE by bs_convert_impl
        if (id < T.n_leaves - 1) {
            const char* nxt = reinterpret_cast<const char*>(T.rec + (size_t)id * REC);
            asm volatile("prefetch.global.L1 [%0];" ::"l"(nxt));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(nxt + 128));
        }
        if (lane == 0) { unsigned* e = stack + sp * 2; e[0] = id; e[1] = near_m; }
        ++sp;
    }
    // two entries of a record at once; every float2 holds (entry a, entry b). VPL == 1 only.
    __device__ __forceinline__ void visit2(int ida, int idb, const float2 hx, const float2 hy, const float2 hz, const float2 hw, const float2* c, unsigned m) {
        const bool in = (m >> lane) & 1u;
        const float2 rx = __fadd2_rn(hx, make_float2(-qx[0], -qx[0])), ry = __fadd2_rn(hy, make_float2(-qy[0], -qy[0])), rz = __fadd2_rn(hz, make_float2(-qz[0], -qz[0]));
        const float2 r2 = __ffma2_rn(rz, rz, __ffma2_rn(ry, ry, __fmul2_rn(rx, rx)));
        const bool fa = in && ida >= 0 && r2.x > hw.x, fb = in && idb >= 0 && r2.y > hw.y;
        const unsigned fma_ = __ballot_sync(0xFFFFFFFFu, fa), fmb = __ballot_sync(0xFFFFFFFFu, fb);
        if (COUNT) { if (lane == 0) cnt[3] += (ida >= 0) + (idb >= 0); if (in) cnt[0] += (ida >= 0) + (idb >= 0); }
        if (fma_ | fmb) {
            float2 inv;
            asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv.x) : "f"(r2.x));
            asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv.y) : "f"(r2.y));
            const float2 ir2 = __fmul2_rn(inv, inv);
            const float2 k = __fmul2_rn(__fmul2_rn(inv, make_float2(INV_4PI, INV_4PI)), ir2);
            // c[0..9] = o1.x o1.y o1.z trM m00 m11 m22 (m01+m10) (m02+m20) (m12+m21)
            const float2 d = __ffma2_rn(c[0], rx, __ffma2_rn(c[1], ry, __ffma2_rn(c[2], rz, c[3])));
            const float2 t1 = __ffma2_rn(c[7], ry, __ffma2_rn(c[8], rz, __fmul2_rn(c[4], rx)));
            const float2 t2 = __ffma2_rn(c[9], rz, __fmul2_rn(c[5], ry));
            const float2 rMr = __ffma2_rn(rx, t1, __ffma2_rn(ry, t2, __fmul2_rn(__fmul2_rn(c[6], rz), rz)));
            const float2 f = __fmul2_rn(k, __ffma2_rn(__fmul2_rn(ir2, make_float2(-3.0f, -3.0f)), rMr, d));
            if (fa) { wn[0] += f.x; if (COUNT) cnt[1]++; }
            if (fb) { wn[0] += f.y; if (COUNT) cnt[1]++; }
        }
        if (ida >= 0) push((unsigned)ida, m & ~fma_);
        if (idb >= 0) push((unsigned)idb, m & ~fmb);
    }
#endif

    // per-voxel traversal of the sub-trees in roots[0..n_roots); wn[] must hold the hoisted far part on entry.
    // All near roots are pushed (and their records prefetched) before the single drain loop starts.
    __device__ void run(const unsigned* valid_m, const unsigned* roots, int n_roots) {
        sp = 0;
        for (int ri = 0; ri < n_roots; ++ri) {  // a node is tested before its type is looked at (aabb_tree.rs:662-670)
            const unsigned rid = roots[ri];
            const float4* c = T.coef + 3 * (size_t)rid;
            visit(rid, __ldg(T.hdr + rid), __ldg(c), __ldg(c + 1), __ldg(c + 2), valid_m);
        }
        while (sp > 0) {
            --sp;
            __syncwarp();
            unsigned m[VPL];
            const unsigned* e = stack + sp * (1 + VPL);
            const unsigned id = e[0];
#pragma unroll
            for (int v = 0; v < VPL; ++v) m[v] = e[1 + v];
            __syncwarp();
            if (id >= T.n_leaves - 1) { exact_leaf(id, m); continue; }
#ifdef BS_PAIRS
            {
                const float4* r = T.rec + (size_t)id * REC;
                const float4 idsf = __ldg(r + 4);
                const int ids[4] = {__float_as_int(idsf.x), __float_as_int(idsf.y), __float_as_int(idsf.z), __float_as_int(idsf.w)};
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    if (ids[2 * p] < 0 && ids[2 * p + 1] < 0) continue;
                    const float4 h0 = __ldg(r + 2 * p), h1 = __ldg(r + 2 * p + 1);
                    float2 c[10];
#pragma unroll
                    for (int k = 0; k < 5; ++k) { const float4 q = __ldg(r + 5 + 5 * p + k); c[2 * k] = make_float2(q.x, q.y); c[2 * k + 1] = make_float2(q.z, q.w); }
                    visit2(ids[2 * p], ids[2 * p + 1], make_float2(h0.x, h0.y), make_float2(h0.z, h0.w), make_float2(h1.x, h1.y), make_float2(h1.z, h1.w), c, m[0]);
                }
            }
#else
            const float4* r = T.rec + (size_t)id * REC;
            const float4 idsf = __ldg(r + 4);
            const int ids[4] = {__float_as_int(idsf.x), __float_as_int(idsf.y), __float_as_int(idsf.z), __float_as_int(idsf.w)};
            // coefficient floats 20..59 = float4 slots 5..14, entry e starts at float 20 + 10 e
            const float4 q5 = __ldg(r + 5), q6 = __ldg(r + 6), q7 = __ldg(r + 7), q8 = __ldg(r + 8), q9 = __ldg(r + 9);
            const float4 q10 = __ldg(r + 10), q11 = __ldg(r + 11), q12 = __ldg(r + 12), q13 = __ldg(r + 13), q14 = __ldg(r + 14);
            if (ids[0] >= 0) visit((unsigned)ids[0], __ldg(r + 0), q5, q6, make_float4(q7.x, q7.y, 0.f, 0.f), m);
            if (ids[1] >= 0) visit((unsigned)ids[1], __ldg(r + 1), make_float4(q7.z, q7.w, q8.x, q8.y), make_float4(q8.z, q8.w, q9.x, q9.y), make_float4(q9.z, q9.w, 0.f, 0.f), m);
            if (ids[2] >= 0) visit((unsigned)ids[2], __ldg(r + 2), q10, q11, make_float4(q12.x, q12.y, 0.f, 0.f), m);
            if (ids[3] >= 0) visit((unsigned)ids[3], __ldg(r + 3), make_float4(q12.z, q12.w, q13.x, q13.y), make_float4(q13.z, q13.w, q14.x, q14.y), make_float4(q14.z, q14.w, 0.f, 0.f), m);
#endif
        }
    }
};

// Brick-level pass (its own kernel, one WARP per brick): walk the top of the tree ONCE per brick, breadth first, one
// lane per frontier node. A node that is far for every voxel of the brick (|p~ - c_B| - rho > 2 r, rho = half
// diagonal of the brick) and at least KAPPA * rho away is "hoisted": evaluated at 27 sample points (3 x 3 x 3
// lattice over the brick, lanes 0..26) and later interpolated tri-quadratically per voxel -- its field varies by
// O((rho/d)^3) across the brick, far below what the 0.2 threshold can see. Everything closer is handed to the
// per-voxel traversal as a list of sub-tree roots, so the per-voxel criterion of the reference (aabb_tree.rs:666)
// still decides there. Lists are built with ballot-ordered compaction: the summation order is deterministic.
constexpr float KAPPA_DEFAULT = 3.0f;  // tri-quadratic interpolation error of a 1/r^2 field at 3 rho: ~1 % of a far contribution (measured: 0 sign changes on every test mesh; 4.0 costs 1.4 ms more)
constexpr int MAX_ROOTS = 96;
constexpr int MAX_HOIST = 640;
constexpr int MAX_FRONT = 96;
constexpr int BP_WARPS = 4;
struct BrickOut { float far[27]; unsigned n_roots; unsigned roots[MAX_ROOTS]; };  // n_roots = 0xFFFFFFFF: no hoisting, start at the tree root

template <bool COUNT>
__global__ void __launch_bounds__(32 * BP_WARPS) k_brick_pass(Tree T, const unsigned long long* __restrict__ keys, size_t n_bricks, const unsigned* __restrict__ brick_list /*bricks that hold work items*/, float vs, float KAPPA, BrickOut* out, unsigned long long* counters) {
    __shared__ unsigned s_front[BP_WARPS][2][MAX_FRONT];
    __shared__ unsigned s_hoist[BP_WARPS][MAX_HOIST];   // (record id << 2 | entry); the root itself is 0xFFFFFFFF
    __shared__ unsigned s_roots[BP_WARPS][MAX_ROOTS];
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5, lt = (1u << lane) - 1;
    const size_t bi = (size_t)blockIdx.x * BP_WARPS + w;
    if (bi >= n_bricks) return;
    const size_t b = brick_list[bi];
    int bx, by, bz; bs_key_brick(keys[b], bx, by, bz);
    const float ox = (float)(bx << 3), oy = (float)(by << 3), oz = (float)(bz << 3);
    const float cx = (ox + 3.5f) * vs, cy = (oy + 3.5f) * vs, cz = (oz + 3.5f) * vs, rho = 6.0621778f * vs, kr = KAPPA * rho;
    unsigned* front[2] = {s_front[w][0], s_front[w][1]};
    unsigned* hoist = s_hoist[w]; unsigned* roots = s_roots[w];
    int nf = 0, nh = 0, nr = 0; bool overflow = false;
    unsigned n_class = 0;
    // 0 hoist, 1 root, 2 descend, 3 none
    auto classify = [&](unsigned id, const float4 h) -> int {
        const float dx = h.x - cx, dy = h.y - cy, dz = h.z - cz;
        const float d = fast_sqrt(dx * dx + dy * dy + dz * dz), br = fast_sqrt(h.w);  // br = beta * radius
        const bool brick_far = (d - rho) > br;
        if (brick_far && d >= kr) return 0;
        if (brick_far || id >= T.n_leaves - 1 || br <= kr) return 1;  // close: the whole sub-tree goes to the per-voxel traversal
        return 2;
    };
    {
        const int c = classify(T.root, __ldg(T.hdr + T.root));
        if (c == 0) { if (lane == 0) hoist[0] = 0xFFFFFFFFu; nh = 1; } else if (c == 1) { if (lane == 0) roots[0] = T.root; nr = 1; } else { if (lane == 0) front[0][0] = T.root; nf = 1; }
        n_class = 1;
    }
    __syncwarp();
    for (int cur = 0; nf > 0 && !overflow; cur ^= 1) {
        int nnext = 0;
        for (int base = 0; base < nf; base += 32) {
            const bool have = base + (int)lane < nf;
            const unsigned id = have ? front[cur][base + lane] : 0;
            int code[4] = {3, 3, 3, 3}; int ids[4] = {-1, -1, -1, -1};
            if (have) {
                const float4* r = T.rec + (size_t)id * REC;
                const float4 idsf = __ldg(r + 4);
                ids[0] = __float_as_int(idsf.x); ids[1] = __float_as_int(idsf.y); ids[2] = __float_as_int(idsf.z); ids[3] = __float_as_int(idsf.w);
#pragma unroll
                for (int k = 0; k < 4; ++k) if (ids[k] >= 0) code[k] = classify((unsigned)ids[k], __ldg(T.hdr + ids[k]));
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // ordered compaction: (round, entry, lane)
                const unsigned mh = __ballot_sync(0xFFFFFFFFu, code[k] == 0), mr = __ballot_sync(0xFFFFFFFFu, code[k] == 1), md = __ballot_sync(0xFFFFFFFFu, code[k] == 2);
                if (COUNT) n_class += __popc(mh | mr | md);
                if (nh + __popc(mh) > MAX_HOIST || nr + __popc(mr) > MAX_ROOTS || nnext + __popc(md) > MAX_FRONT) { overflow = true; break; }
                if (code[k] == 0) hoist[nh + __popc(mh & lt)] = (unsigned)ids[k];
                if (code[k] == 1) roots[nr + __popc(mr & lt)] = (unsigned)ids[k];
                if (code[k] == 2) front[cur ^ 1][nnext + __popc(md & lt)] = (unsigned)ids[k];
                nh += __popc(mh); nr += __popc(mr); nnext += __popc(md);
            }
            if (overflow) break;
        }
        __syncwarp();
        nf = nnext;
    }
    BrickOut* o = out + b;
    if (overflow) {  // pathological tree: plain per-voxel traversal from the root, nothing hoisted
        if (lane < 27) o->far[lane] = 0.f;
        if (lane == 0) o->n_roots = 0xFFFFFFFFu;
        return;
    }
    // hoisted nodes at the 27 sample points
    const unsigned si = lane < 27 ? lane / 9 : 0, sj = lane < 27 ? (lane / 3) % 3 : 0, sk = lane < 27 ? lane % 3 : 0;
    const float sx = (ox + 3.5f * si) * vs, sy = (oy + 3.5f * sj) * vs, sz = (oz + 3.5f * sk) * vs;
    float S = 0.f;
    for (int i = 0; i < nh; ++i) {
        const unsigned e = hoist[i];
        const unsigned nid = e == 0xFFFFFFFFu ? T.root : e;  // warp-uniform: broadcast loads
        const float4 h = __ldg(T.hdr + nid);
        const float4 a0 = __ldg(T.coef + 3 * (size_t)nid), a1 = __ldg(T.coef + 3 * (size_t)nid + 1), a2 = __ldg(T.coef + 3 * (size_t)nid + 2);
        const float c[10] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y};
        const float rx = h.x - sx, ry = h.y - sy, rz = h.z - sz;
        const float r2 = fmaf(rz, rz, fmaf(ry, ry, rx * rx));
        float inv_r;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv_r) : "f"(r2));
        const float ir2 = inv_r * inv_r, k = (INV_4PI * inv_r) * ir2;
        const float dd = fmaf(c[0], rx, fmaf(c[1], ry, fmaf(c[2], rz, c[3])));
        const float t1 = fmaf(c[7], ry, fmaf(c[8], rz, c[4] * rx)), t2 = fmaf(c[9], rz, c[5] * ry);
        const float rMr = fmaf(rx, t1, fmaf(ry, t2, (c[6] * rz) * rz));
        S += k * fmaf(-3.0f * ir2, rMr, dd);
    }
    if (lane < 27) o->far[lane] = S;
    if (lane == 0) o->n_roots = (unsigned)nr;
    for (int i = lane; i < nr; i += 32) o->roots[i] = roots[i];
    if (COUNT && lane == 0) { atomicAdd(counters + 6, (unsigned long long)n_class); atomicAdd(counters + 7, (unsigned long long)nh); atomicAdd(counters + 8, (unsigned long long)nr); }
}

// 9-bit Morton position inside a brick -> leaf offset x<<6 | y<<3 | z (bits 2,5,8 -> x; 1,4,7 -> y; 0,3,6 -> z)
__device__ __forceinline__ unsigned demorton9(unsigned p) {
    const unsigned x = ((p >> 2) & 1) | ((p >> 4) & 2) | ((p >> 6) & 4);
    const unsigned y = ((p >> 1) & 1) | ((p >> 3) & 2) | ((p >> 5) & 4);
    const unsigned z = (p & 1) | ((p >> 2) & 2) | ((p >> 4) & 4);
    return (x << 6) | (y << 3) | z;
}

// Active masks + the list of work items for the sign kernel: one item = 32*VPL Morton-adjacent active voxels of one brick.
// The distance pass (k_eval, bs_convert.cu) leaves the minimum SQUARED distance in every touched voxel: the root is taken
// here, once per voxel instead of once per point-triangle pair (sqrt is monotone: root of the minimum == minimum of the roots,
// bit for bit).
__global__ void k_masks(float* __restrict__ values, size_t n_bricks, unsigned long long* masks, unsigned* n_chunks, int per_chunk, unsigned long long* n_active) {
    const unsigned lane = threadIdx.x & 31;
    const size_t b = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    if (b >= n_bricks) return;
    float* bv = values + b * 512;
    unsigned lo = 0, hi = 0, cnt = 0;
    for (int r = 0; r < 16; ++r) {
        const float v2 = bv[r * 32 + lane];
        const bool act = __float_as_uint(v2) != BS_UDF_SENTINEL_BITS;
        if (act) bv[r * 32 + lane] = __fsqrt_rn(v2);
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, act);
        if ((int)lane == (r >> 1)) { if (r & 1) hi = bal; else lo = bal; }
        cnt += __popc(bal);
    }
    if (lane < 8) masks[b * 8 + lane] = (unsigned long long)lo | ((unsigned long long)hi << 32);
    if (lane == 0) { n_chunks[b] = (cnt + per_chunk - 1) / per_chunk; if (cnt) atomicAdd(n_active, (unsigned long long)cnt); }
}
// items are laid out in `order` (bricks under the densest triangles first: their items run longest, so they must not
// start last); brick_off[b] = first item of brick b
__global__ void k_order_chunks(const unsigned* __restrict__ order, const unsigned* __restrict__ n_chunks, size_t n_bricks, unsigned* ordered) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n_bricks) ordered[i] = n_chunks[order ? order[i] : i];
    if (i == n_bricks) ordered[i] = 0;
}
__global__ void k_items(const unsigned* __restrict__ order, const unsigned* __restrict__ off, size_t n_bricks, unsigned* item_brick, unsigned* brick_off) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n_bricks) return;
    const unsigned b = order ? order[i] : (unsigned)i;
    brick_off[b] = off[i];
    for (unsigned k = off[i]; k < off[i + 1]; ++k) item_brick[k] = b;
}
// bricks with at least one work item (with sign propagation most bricks have none)
__global__ void k_item_bricks(const unsigned* __restrict__ n_chunks, size_t n_bricks, unsigned* list, unsigned* count) {
    const size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const bool has = b < n_bricks && n_chunks[b] != 0;
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, has), lane = threadIdx.x & 31;
    unsigned base = 0;
    if (lane == 0 && bal) base = atomicAdd(count, (unsigned)__popc(bal));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (has) list[base + __popc(bal & ((1u << lane) - 1))] = (unsigned)b;
}
__global__ void k_touch_keys(const unsigned long long* __restrict__ touches, const unsigned long long* __restrict__ total, size_t n, int shift, int buckets, unsigned* key, unsigned* idx) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long t = touches[i], m = *total / n + 1;
    // Bricks keep their spatial (visit) order, which is what the caches like, unless they sit under unusually dense
    // triangles: those move to the front in buckets of doubling density (their items run long and must not start last;
    // on a brick slab whose dense region comes last in key order that tail was 3 ms of an 11 ms stage).
    // Bits 8..: heavy bricks (>= 2^shift touches), densest first; bits 0..7: log2 bucket of t / mean.
    unsigned k = 0;
    if (buckets && t >= 2 * m) k = 1u + (unsigned)(63 - __clzll((long long)(t / m)));  // (single GPU: the tail hides behind 30 ms of other items, plain order is 0.3 ms faster)
    const unsigned long long hv = t >> shift;
    key[i] = (hv ? ((hv > 0xFFFFFFull ? 0xFFFFFFu : (unsigned)hv) << 8) : 0u) | k;
    idx[i] = (unsigned)i;
}
// ---- heavy bricks --------------------------------------------------------------------------------------------------------
// A brick under thousands of sliver triangles (the pole of a UV sphere) makes every one of its items walk ~10^4 leaves:
// milliseconds for a single warp, which is the floor of the whole stage once the bricks are spread over 4-8 GPUs.
// Those items are split by TRIANGLES: the brick's root list is expanded breadth first into up to HEAVY_CAP small
// sub-trees, HEAVY_REPL warps per item each walk every HEAVY_REPL-th of them, and k_sign_finish adds the partial sums in
// a fixed order (deterministic) and sets the sign. A node that would have been accepted as far may be replaced by its
// descendants here, each still tested with the per-voxel criterion: never less accurate than the unsplit walk.
constexpr int HEAVY_CAP = 1024, HEAVY_REPL = 16;  // <= HEAVY_CAP / HEAVY_REPL = 64 roots per replica (MAX_ROOTS = 96)
struct HeavyRoots { unsigned n; unsigned ids[HEAVY_CAP]; };
struct HeavySplit { unsigned repl, n_hitems; const HeavyRoots* roots; const unsigned* slot_of_brick; float* partial; unsigned short* offs; };

// one warp per heavy brick. Entries are expanded by SIZE (leaves under a Karras node i: |i - other[i]| + 1): every round
// replaces the internal entries holding more than 1/256 of the list's leaves by the entries of their records, so the
// fan of slivers ends up in a few hundred pieces of similar size that the replicas share round-robin.
__global__ void __launch_bounds__(128) k_expand_roots(Tree T, const int* __restrict__ other, const unsigned* __restrict__ heavy_bricks, unsigned n_heavy, const BrickOut* __restrict__ brick_out, const unsigned* __restrict__ n_chunks, HeavyRoots* out) {
    __shared__ unsigned s_l[4][2][HEAVY_CAP];
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5, h = blockIdx.x * 4 + w;
    if (h >= n_heavy) return;
    if (n_chunks[heavy_bricks[h]] == 0) { if (lane == 0) out[h].n = 0; return; }  // no work items: the brick pass skipped this brick
    const BrickOut* bo = brick_out + heavy_bricks[h];
    unsigned n = bo->n_roots;
    if (n == 0xFFFFFFFFu) { if (lane == 0) s_l[w][0][0] = T.root; n = 1; }
    else for (unsigned i = lane; i < n; i += 32) s_l[w][0][i] = bo->roots[i];
    __syncwarp();
    auto size_of = [&](unsigned id) -> unsigned { return id >= T.n_leaves - 1 ? 1u : (unsigned)abs((int)id - other[id]) + 1u; };
    int cur = 0;
    for (int round = 0; round < 24; ++round) {
        unsigned leaves = 0;
        for (unsigned i = lane; i < n; i += 32) leaves += size_of(s_l[w][cur][i]);
#pragma unroll
        for (int o = 16; o; o >>= 1) leaves += __shfl_xor_sync(0xFFFFFFFFu, leaves, o);
        const unsigned thr = max(16u, leaves / 256u);
        unsigned total = 0, big = 0;
        for (unsigned base = 0; base < n; base += 32) {
            const unsigned i = base + lane;
            unsigned c = 0;
            if (i < n) {
                const unsigned id = s_l[w][cur][i];
                if (size_of(id) <= thr) c = 1;
                else { const float4 idsf = __ldg(T.rec + (size_t)id * REC + 4); c = 2 + (__float_as_int(idsf.z) >= 0) + (__float_as_int(idsf.w) >= 0); big = 1; }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
            total += c;
        }
        if (!__any_sync(0xFFFFFFFFu, big) || total > (unsigned)HEAVY_CAP) break;
        unsigned wr = 0;
        for (unsigned base = 0; base < n; base += 32) {
            const unsigned i = base + lane;
            int ids[4] = {-1, -1, -1, -1}; unsigned c = 0;
            if (i < n) {
                const unsigned id = s_l[w][cur][i];
                if (size_of(id) <= thr) { ids[0] = (int)id; c = 1; }
                else { const float4 idsf = __ldg(T.rec + (size_t)id * REC + 4); ids[0] = __float_as_int(idsf.x); ids[1] = __float_as_int(idsf.y); ids[2] = __float_as_int(idsf.z); ids[3] = __float_as_int(idsf.w); c = 2 + (ids[2] >= 0) + (ids[3] >= 0); }
            }
            unsigned inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if ((int)lane >= o) inc += t; }
            unsigned p = wr + inc - c;
            for (int k = 0; k < 4; ++k) if (ids[k] >= 0) s_l[w][cur ^ 1][p++] = (unsigned)ids[k];
            wr += __shfl_sync(0xFFFFFFFFu, inc, 31);
        }
        __syncwarp();
        cur ^= 1; n = wr;
    }
    HeavyRoots* o = out + h;
    if (lane == 0) o->n = n;
    for (unsigned i = lane; i < n; i += 32) o->ids[i] = s_l[w][cur][i];
}

// one warp per heavy item: partial sums in replica order, then the sign (mesh_to_volume.rs:266-271)
template <int VPL>
__global__ void k_sign_finish(float* values, const unsigned* __restrict__ item_brick, unsigned first_item, unsigned n_items, unsigned repl, const float* __restrict__ partial, const unsigned short* __restrict__ offs) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= (size_t)n_items * 32 * VPL) return;
    const size_t it = t / (32 * VPL); const unsigned slot = (unsigned)(t % (32 * VPL));
    const unsigned off = offs[t];
    if (off == 0xFFFFu) return;
    float wn = 0.f;
    for (unsigned r = 0; r < repl; ++r) wn += partial[(it * repl + r) * (32 * VPL) + slot];
    float* bv = values + (size_t)item_brick[first_item + it] * 512;
    const float d = bv[off];
    bv[off] = (wn < 0.2f) ? copysignf(d, 1.0f) : copysignf(d, -1.0f);
}
__global__ void k_heavy_list(const unsigned* __restrict__ order, const unsigned* __restrict__ keys_sorted_desc, size_t n, unsigned* heavy_bricks, unsigned* slot_of_brick, unsigned* n_heavy) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if ((keys_sorted_desc[i] >> 8) > 0) { heavy_bricks[i] = order[i]; slot_of_brick[order[i]] = (unsigned)i; atomicMax(n_heavy, (unsigned)i + 1); }  // heavy bricks are the first entries of `order`
}

// One WARP per work item (32*VPL Morton-adjacent active voxels of one brick): a heavy brick (e.g. at the pole of a UV
// sphere, where thousands of sliver triangles are "near") is spread over many warps / SMs instead of serialising a CTA.
template <bool COUNT, int VPL>
#ifndef BS_SIGN_MINB
#define BS_SIGN_MINB 8
#endif
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK, BS_SIGN_MINB) k_sign(Tree T, float* values, const unsigned long long* __restrict__ masks, unsigned n_items,
                                                               const unsigned* __restrict__ item_brick, const unsigned* __restrict__ chunk_off,
                                                               const unsigned long long* __restrict__ keys, float vs, const BrickOut* __restrict__ brick_out, unsigned long long* counters,
                                                               HeavySplit H) {
    __shared__ unsigned s_stack[WARPS_PER_BLOCK][STACK * (1 + VPL)];
    __shared__ unsigned short s_list[WARPS_PER_BLOCK][32 * VPL];
    __shared__ float s_far[WARPS_PER_BLOCK][27];
    __shared__ unsigned s_roots[WARPS_PER_BLOCK][MAX_ROOTS];
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // The first H.n_hitems items belong to heavy bricks: H.repl warps each, replica r walks every H.repl-th entry of the
    // brick's expanded root list and leaves a partial sum for k_sign_finish. All other items: one warp. One launch, so
    // the ordinary items fill the SMs while the long replicas run.
    const unsigned wi = blockIdx.x * WARPS_PER_BLOCK + w;
    const unsigned nhw = H.n_hitems * H.repl;
    const bool heavy = wi < nhw;
    const unsigned item = heavy ? wi / H.repl : H.n_hitems + (wi - nhw), rep = heavy ? wi % H.repl : 0;
    if (item >= n_items) return;
    const size_t b = item_brick[item];
    const unsigned chunk = item - chunk_off[b];
    float* bv = values + b * 512;
    int bx, by, bz; bs_key_brick(keys[b], bx, by, bz);
    int n_roots;
    {   // result of the brick-level pass: hoisted far field at 27 samples + sub-tree roots
        const BrickOut* bo = brick_out + b;
        const unsigned nr = bo->n_roots;
        if (lane < 27) s_far[w][lane] = bo->far[lane];
        if (heavy) {
            const HeavyRoots* hr = H.roots + H.slot_of_brick[b];
            const unsigned n_all = hr->n;
            n_roots = 0;
            for (unsigned i = rep + lane * H.repl; i < n_all; i += 32 * H.repl) s_roots[w][(i - rep) / H.repl] = hr->ids[i];
            n_roots = (int)((n_all > rep) ? (n_all - rep + H.repl - 1) / H.repl : 0);
        }
        else if (nr == 0xFFFFFFFFu) { if (lane == 0) s_roots[w][0] = T.root; n_roots = 1; }
        else {
            for (unsigned i = lane; i < nr; i += 32) {
                const unsigned rid = bo->roots[i];
                s_roots[w][i] = rid;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(T.hdr + rid));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(T.coef + 3 * (size_t)rid));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(T.coef + 3 * (size_t)rid + 2));
            }
            n_roots = (int)nr;
        }
    }
    // this item's voxels: entries [chunk*32*VPL, +32*VPL) of the brick's active voxels in Morton order
    unsigned cnt = 0;  // active voxels of the brick seen so far
    const unsigned first = chunk * 32 * VPL;
    const unsigned long long* mk = masks + b * 8;
    for (int r = 0; r < 16; ++r) {
        const unsigned off = demorton9(r * 32 + lane);
        const bool act = (mk[off >> 6] >> (off & 63)) & 1;
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, act);
        const unsigned p = cnt + __popc(bal & ((1u << lane) - 1));
        if (act && p >= first && p < first + 32 * VPL) s_list[w][p - first] = (unsigned short)off;
        cnt += __popc(bal);
    }
    __syncwarp();
    const unsigned n_here = min(cnt - first, (unsigned)(32 * VPL));
    {
        unsigned c3[4] = {0, 0, 0, 0};
        WarpWinding<COUNT, VPL> W{T};
        W.stack = s_stack[w]; W.cnt = c3; W.lane = lane;
        unsigned off[VPL], valid_m[VPL]; bool valid[VPL];
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const unsigned i = v * 32 + lane;
            valid[v] = i < n_here;
            valid_m[v] = __ballot_sync(0xFFFFFFFFu, valid[v]);
            off[v] = s_list[w][valid[v] ? i : n_here - 1];
            W.qx[v] = __fmul_rn((float)((bx << 3) + (int)(off[v] >> 6)), vs);
            W.qy[v] = __fmul_rn((float)((by << 3) + (int)((off[v] >> 3) & 7)), vs);
            W.qz[v] = __fmul_rn((float)((bz << 3) + (int)(off[v] & 7)), vs);
            // hoisted far part: tri-quadratic Lagrange interpolation on the nodes {0, 3.5, 7} per axis
            float L[3][3];
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                const float t = (float)(ax == 0 ? (off[v] >> 6) : (ax == 1 ? ((off[v] >> 3) & 7) : (off[v] & 7)));
                L[ax][0] = (t - 3.5f) * (t - 7.0f) * (1.0f / 24.5f); L[ax][1] = t * (7.0f - t) * (1.0f / 12.25f); L[ax][2] = t * (t - 3.5f) * (1.0f / 24.5f);
            }
            float acc = 0.f;
#pragma unroll
            for (int i2 = 0; i2 < 3; ++i2)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const float wij = L[0][i2] * L[1][j];
                    acc = fmaf(wij, fmaf(L[2][0], s_far[w][i2 * 9 + j * 3], fmaf(L[2][1], s_far[w][i2 * 9 + j * 3 + 1], L[2][2] * s_far[w][i2 * 9 + j * 3 + 2])), acc);
                }
            W.wn[v] = (heavy && rep) ? 0.f : acc;  // the hoisted part is counted once
        }
        __syncwarp();
        W.run(valid_m, s_roots[w], n_roots);
        if (heavy) {  // partial sums and (once) the voxel offsets, for the finish kernel
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                H.partial[((size_t)item * H.repl + rep) * (32 * VPL) + v * 32 + lane] = W.wn[v];
                if (rep == 0) H.offs[(size_t)item * (32 * VPL) + v * 32 + lane] = valid[v] ? (unsigned short)off[v] : (unsigned short)0xFFFFu;
            }
            if (COUNT) { if (rep == 0) { for (int v = 0; v < VPL; ++v) if (valid[v]) atomicAdd(counters + 3, 1ull); } atomicAdd(counters, (unsigned long long)c3[0]); atomicAdd(counters + 1, (unsigned long long)c3[1]); atomicAdd(counters + 2, (unsigned long long)c3[2]); atomicAdd(counters + 4, (unsigned long long)c3[3]); atomicAdd(counters + 5, (unsigned long long)(lane == 0)); }
            return;
        }
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            if (!valid[v]) continue;
            const float d = bv[off[v]];
            bv[off[v]] = (W.wn[v] < 0.2f) ? copysignf(d, 1.0f) : copysignf(d, -1.0f);  // mesh_to_volume.rs:266-271
            if (COUNT) atomicAdd(counters + 3, 1ull);
        }
        if (COUNT) { atomicAdd(counters, (unsigned long long)c3[0]); atomicAdd(counters + 1, (unsigned long long)c3[1]); atomicAdd(counters + 2, (unsigned long long)c3[2]); atomicAdd(counters + 4, (unsigned long long)c3[3]); atomicAdd(counters + 5, (unsigned long long)(lane == 0)); }
    }
}

}  // namespace

#ifndef BS_VPL
#define BS_VPL 1
#endif

namespace {
struct TreeBuffers { float4 *sorted = nullptr, *hdr = nullptr, *coef = nullptr, *rec = nullptr; int* other = nullptr; };
void free_tree(bs_context* ctx, TreeBuffers& B) { bs_free(ctx, B.sorted); bs_free(ctx, B.hdr); bs_free(ctx, B.coef); bs_free(ctx, B.rec); bs_free(ctx, B.other); B = TreeBuffers(); }
// LBVH over the triangles: Morton order, Karras hierarchy, moments, traversal records
bs_status build_tree(bs_context* ctx, const float* d_tris, size_t n_tris, Tree& T, TreeBuffers& B) {
    cudaStream_t st = ctx->stream;
    int* d_bounds = nullptr; unsigned long long *d_codes = nullptr, *d_codes2 = nullptr; unsigned *d_ids = nullptr, *d_ids2 = nullptr;
    BS_TRY(bs_alloc(ctx, &d_bounds, 6));
    const int init[6] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF, (int)0x80000000, (int)0x80000000, (int)0x80000000};
    BS_CUDA(ctx, cudaMemcpyAsync(d_bounds, init, sizeof(init), cudaMemcpyHostToDevice, st));
    BS_TRY(bs_alloc(ctx, &d_codes, n_tris)); BS_TRY(bs_alloc(ctx, &d_codes2, n_tris));
    BS_TRY(bs_alloc(ctx, &d_ids, n_tris)); BS_TRY(bs_alloc(ctx, &d_ids2, n_tris));
    bs_count_launch(), k_centroid_bounds<<<(unsigned)std::min<size_t>(bs_blocks(n_tris, 256), (size_t)ctx->sm_count * 64), 256, 0, st>>>(d_tris, n_tris, d_bounds);
    bs_count_launch(), k_morton<<<bs_blocks(n_tris, 256), 256, 0, st>>>(d_tris, n_tris, d_bounds, d_codes, d_ids);
    void* d_tmp = nullptr; size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_codes, d_codes2, d_ids, d_ids2, n_tris, 0, 48, st);
    BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp_bytes));
    cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_codes, d_codes2, d_ids, d_ids2, n_tris, 0, 48, st);
    bs_free(ctx, d_tmp); bs_free(ctx, d_codes); bs_free(ctx, d_ids); bs_free(ctx, d_bounds);
    bs_mark(ctx, "bvh_sort_ms");
    const int n = (int)((n_tris + LEAF - 1) / LEAF);
    const int n_nodes = 2 * n - 1;
    int *d_left = nullptr, *d_right = nullptr, *d_parent = nullptr; unsigned* d_flags = nullptr; Raw* d_raw = nullptr;
    BS_TRY(bs_alloc(ctx, &d_left, (size_t)n)); BS_TRY(bs_alloc(ctx, &d_right, (size_t)n)); BS_TRY(bs_alloc(ctx, &d_parent, (size_t)n_nodes));
    BS_TRY(bs_alloc(ctx, &d_flags, (size_t)n)); BS_TRY(bs_alloc(ctx, &B.other, (size_t)n));
    BS_TRY(bs_alloc(ctx, &B.sorted, (size_t)n * LEAF * 3));
    BS_TRY(bs_alloc(ctx, &B.hdr, (size_t)n_nodes)); BS_TRY(bs_alloc(ctx, &B.coef, (size_t)n_nodes * 3));
    BS_TRY(bs_alloc(ctx, &B.rec, (size_t)n * REC));
    BS_TRY(bs_alloc(ctx, &d_raw, (size_t)n_nodes));
    BS_CUDA(ctx, cudaMemsetAsync(d_flags, 0, (size_t)n * sizeof(unsigned), st));
    if (n > 1) bs_count_launch(), k_karras<<<bs_blocks((size_t)n - 1, 256), 256, 0, st>>>(d_codes2, n, d_left, d_right, d_parent, B.other);
    bs_mark(ctx, "bvh_tree_ms");
    bs_count_launch(), k_leaves_and_climb<<<bs_blocks((size_t)n, CLIMB_TPB), CLIMB_TPB, 0, st>>>(d_tris, d_ids2, n_tris, B.sorted, d_raw, d_left, d_right, d_parent, B.other, d_flags, n, B.hdr, B.coef);
    bs_mark(ctx, "bvh_moments_ms");
    if (n > 1) bs_count_launch(), k_records<<<bs_blocks(((size_t)n - 1) * 4, 256), 256, 0, st>>>(d_left, d_right, B.hdr, B.coef, n, B.rec);
    bs_free(ctx, d_raw); bs_free(ctx, d_ids2); bs_free(ctx, d_codes2); bs_free(ctx, d_left); bs_free(ctx, d_right); bs_free(ctx, d_parent); bs_free(ctx, d_flags);
    T.hdr = B.hdr; T.coef = B.coef; T.rec = B.rec; T.tris = B.sorted; T.n_leaves = (unsigned)n; T.root = 0u; T.derr = ctx->d_err;
    bs_mark(ctx, "bvh_records_ms");
    return BS_OK;
}
}  // namespace

bs_status bs_sign_impl(bs_context* ctx, const float* d_tris, size_t n_tris, bs_volume* vol, const unsigned long long* d_touches, const unsigned long long* d_blk) {
    cudaStream_t st = ctx->stream;
    if (n_tris >= (1ull << 30)) return bs_fail(ctx, BS_ERR_RANGE, "too many triangles");
    const size_t nb = vol->n_bricks;
    if (nb == 0) return BS_OK;
    // closed mesh (tested by bs_convert_impl, bs_signprop.cu): only one voxel per connected band component is evaluated,
    // the others copy its sign
    bool prop = ctx->sign_propagation && ctx->mesh_closed && d_blk != nullptr;
    void* d_tmp = nullptr; size_t tmp_bytes = 0;
    unsigned *d_nchunks = nullptr, *d_ordered = nullptr, *d_off = nullptr, *d_chunk_off = nullptr, *d_item_brick = nullptr, *d_order = nullptr; unsigned n_items = 0;
    BS_TRY(bs_alloc(ctx, &d_nchunks, nb)); BS_TRY(bs_alloc(ctx, &d_ordered, nb + 1)); BS_TRY(bs_alloc(ctx, &d_off, nb + 1)); BS_TRY(bs_alloc(ctx, &d_chunk_off, nb));
    unsigned long long* d_sc = nullptr; unsigned long long h_sc[2] = {0, 0};  // active voxels, evaluated representatives
    BS_TRY(bs_alloc(ctx, &d_sc, 2));
    BS_CUDA(ctx, cudaMemsetAsync(d_sc, 0, 2 * sizeof(unsigned long long), st));
    bs_count_launch(), k_masks<<<bs_blocks(nb * 32, 256), 256, 0, st>>>(vol->values, nb, vol->masks, d_nchunks, 32 * BS_VPL, d_sc);
    bs_sign_components C; memset(&C, 0, sizeof(C));
    bool brute = false;
    if (prop) {  // work items are formed from the component representatives instead of all active voxels
        BS_TRY(bs_sign_components_impl(ctx, vol, d_blk, &C, d_nchunks, 32 * BS_VPL, d_sc + 1));
        prop = C.ok;
        if (prop) {
            BS_TRY(bs_fetch(ctx, h_sc, d_sc, sizeof(h_sc)));
            BS_TRY(bs_sync(ctx));
            brute = h_sc[1] <= (unsigned long long)bs_sign_brute_max() && !ctx->count_work && !getenv("BSHARK_NO_BRUTE");
        }
        bs_mark(ctx, "sign_components_ms");
    }
    unsigned n_heavy = 0, n_hitems = 0;
    if (brute) {
        BS_TRY(bs_sign_brute_impl(ctx, d_tris, n_tris, vol, &C, (unsigned)h_sc[1]));
        bs_mark(ctx, "sign_brute_ms");
    } else {
        Tree T; TreeBuffers TB;
        BS_TRY(build_tree(ctx, d_tris, n_tris, T, TB));
        const unsigned long long* item_masks = prop ? C.seed : vol->masks;
        unsigned *d_heavy = nullptr, *d_slot = nullptr, *d_nheavy = nullptr;
        if (d_touches) {  // heaviest bricks first; bricks with >= 2^shift touching sub-triangle boxes are "heavy" (split by triangles)
            int shift = 11;
            // On one GPU with every voxel traversed the long items simply start first and hide behind the rest (splitting them
            // costs 1.6 ms of extra launches and refinement on config 5); on a brick slab, or when only the representatives
            // are traversed, they ARE the stage time, so those runs split.
            bool split = vol->owned != nullptr || prop;
            if (const char* e = getenv("BSHARK_HEAVY_SHIFT")) { shift = atoi(e); split = true; }  // tests: force the heavy path on small meshes
            unsigned *d_k = nullptr, *d_k2 = nullptr, *d_i = nullptr;
            BS_TRY(bs_alloc(ctx, &d_k, nb)); BS_TRY(bs_alloc(ctx, &d_k2, nb)); BS_TRY(bs_alloc(ctx, &d_i, nb)); BS_TRY(bs_alloc(ctx, &d_order, nb));
            BS_TRY(bs_alloc(ctx, &d_heavy, nb)); BS_TRY(bs_alloc(ctx, &d_slot, nb)); BS_TRY(bs_alloc(ctx, &d_nheavy, 1));
            BS_CUDA(ctx, cudaMemsetAsync(d_nheavy, 0, sizeof(unsigned), st));
            unsigned long long* d_tot = nullptr;
            BS_TRY(bs_alloc(ctx, &d_tot, 1));
            tmp_bytes = 0;
            cub::DeviceReduce::Sum(nullptr, tmp_bytes, d_touches, d_tot, (int)nb, st);
            BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp_bytes));
            cub::DeviceReduce::Sum(d_tmp, tmp_bytes, d_touches, d_tot, (int)nb, st);
            bs_free(ctx, d_tmp);
            bs_count_launch(), k_touch_keys<<<bs_blocks(nb, 256), 256, 0, st>>>(d_touches, d_tot, nb, shift, split ? 1 : 0, d_k, d_i);
            bs_free(ctx, d_tot);
            tmp_bytes = 0;
            cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, d_k, d_k2, d_i, d_order, nb, 0, 32, st);
            BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp_bytes));
            cub::DeviceRadixSort::SortPairsDescending(d_tmp, tmp_bytes, d_k, d_k2, d_i, d_order, nb, 0, 32, st);
            if (split) bs_count_launch(), k_heavy_list<<<bs_blocks(nb, 256), 256, 0, st>>>(d_order, d_k2, nb, d_heavy, d_slot, d_nheavy);
            BS_TRY(bs_fetch(ctx, &n_heavy, d_nheavy, sizeof(unsigned)));
            bs_free(ctx, d_tmp); bs_free(ctx, d_k); bs_free(ctx, d_k2); bs_free(ctx, d_i); bs_free(ctx, d_nheavy);
        }
        unsigned *d_blist = nullptr, *d_nblist = nullptr; unsigned n_blist = 0;
        BS_TRY(bs_alloc(ctx, &d_blist, nb)); BS_TRY(bs_alloc(ctx, &d_nblist, 1));
        BS_CUDA(ctx, cudaMemsetAsync(d_nblist, 0, sizeof(unsigned), st));
        bs_count_launch(), k_item_bricks<<<bs_blocks(nb, 256), 256, 0, st>>>(d_nchunks, nb, d_blist, d_nblist);
        BS_TRY(bs_fetch(ctx, &n_blist, d_nblist, sizeof(unsigned)));
        bs_count_launch(), k_order_chunks<<<bs_blocks(nb + 1, 256), 256, 0, st>>>(d_order, d_nchunks, nb, d_ordered);
        tmp_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_ordered, d_off, nb + 1, st);
        BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp_bytes));
        cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_ordered, d_off, nb + 1, st);
        BS_TRY(bs_fetch(ctx, &n_items, d_off + nb, sizeof(unsigned)));
        BS_TRY(bs_fetch(ctx, h_sc, d_sc, sizeof(h_sc)));
        BS_TRY(bs_sync(ctx));  // n_heavy has arrived too
        if (n_heavy) {  // the heavy bricks head `order`: their items are items [0, n_hitems)
            BS_TRY(bs_fetch(ctx, &n_hitems, d_off + n_heavy, sizeof(unsigned)));
            BS_TRY(bs_sync(ctx));
        }
        bs_free(ctx, d_tmp);
        BS_TRY(bs_alloc(ctx, &d_item_brick, n_items));
        bs_count_launch(), k_items<<<bs_blocks(nb, 256), 256, 0, st>>>(d_order, d_off, nb, d_item_brick, d_chunk_off);
        BrickOut* d_bo = nullptr;
        float kappa = KAPPA_DEFAULT;
        if (const char* e = getenv("BSHARK_KAPPA")) kappa = (float)atof(e);  // tuning knob for experiments only
        BS_TRY(bs_alloc(ctx, &d_bo, nb));
        HeavyRoots* d_hr = nullptr; float* d_partial = nullptr; unsigned short* d_offs = nullptr;
        if (n_hitems) {
            BS_TRY(bs_alloc(ctx, &d_hr, n_heavy)); BS_TRY(bs_alloc(ctx, &d_partial, (size_t)n_hitems * HEAVY_REPL * 32 * BS_VPL)); BS_TRY(bs_alloc(ctx, &d_offs, (size_t)n_hitems * 32 * BS_VPL));
        }
        const HeavySplit H{(unsigned)HEAVY_REPL, n_hitems, d_hr, d_slot, d_partial, d_offs};
        const size_t n_warps = (size_t)n_hitems * HEAVY_REPL + (n_items - n_hitems);
        const unsigned blocks = (unsigned)((n_warps + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
        unsigned long long* d_cnt = nullptr;
        if (ctx->count_work) { BS_TRY(bs_alloc(ctx, &d_cnt, 9)); BS_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, 72, st)); }
        if (n_blist) {
            if (ctx->count_work) bs_count_launch(), k_brick_pass<true><<<bs_blocks(n_blist, BP_WARPS), 32 * BP_WARPS, 0, st>>>(T, vol->keys, n_blist, d_blist, vol->voxel_size, kappa, d_bo, d_cnt);
            else bs_count_launch(), k_brick_pass<false><<<bs_blocks(n_blist, BP_WARPS), 32 * BP_WARPS, 0, st>>>(T, vol->keys, n_blist, d_blist, vol->voxel_size, kappa, d_bo, nullptr);
        }
        if (!ctx->count_work) bs_mark(ctx, "sign_brick_pass_ms");
        if (n_hitems) bs_count_launch(), k_expand_roots<<<bs_blocks(n_heavy, 4), 128, 0, st>>>(T, TB.other, d_heavy, n_heavy, d_bo, d_nchunks, d_hr);
        if (n_items) {
            if (ctx->count_work) bs_count_launch(), k_sign<true, BS_VPL><<<blocks, 32 * WARPS_PER_BLOCK, 0, st>>>(T, vol->values, item_masks, n_items, d_item_brick, d_chunk_off, vol->keys, vol->voxel_size, d_bo, d_cnt, H);
            else bs_count_launch(), k_sign<false, BS_VPL><<<blocks, 32 * WARPS_PER_BLOCK, 0, st>>>(T, vol->values, item_masks, n_items, d_item_brick, d_chunk_off, vol->keys, vol->voxel_size, d_bo, nullptr, H);
        }
        if (n_hitems) bs_count_launch(), k_sign_finish<BS_VPL><<<bs_blocks((size_t)n_hitems * 32 * BS_VPL, 256), 256, 0, st>>>(vol->values, d_item_brick, 0u, n_hitems, (unsigned)HEAVY_REPL, d_partial, d_offs);
        if (ctx->count_work) {
            unsigned long long h_cnt[9];
            BS_TRY(bs_fetch(ctx, h_cnt, d_cnt, 72));
            BS_TRY(bs_sync(ctx));
            bs_free(ctx, d_cnt);
            for (int i = 0; i < 9; ++i) ctx->fwn_counts[i] = (double)h_cnt[i];
        }
        bs_free(ctx, d_hr); bs_free(ctx, d_partial); bs_free(ctx, d_offs); bs_free(ctx, d_heavy); bs_free(ctx, d_slot);
        bs_free(ctx, d_bo); bs_free(ctx, d_item_brick); bs_free(ctx, d_blist); bs_free(ctx, d_nblist);
        free_tree(ctx, TB);
        bs_mark(ctx, "sign_ms");
    }
    if (prop) { BS_TRY(bs_sign_broadcast_impl(ctx, vol, &C)); bs_mark(ctx, "sign_broadcast_ms"); }
    bs_sign_components_free(ctx, &C);
    bs_free(ctx, d_sc); bs_free(ctx, d_nchunks); bs_free(ctx, d_ordered); bs_free(ctx, d_off); bs_free(ctx, d_order); bs_free(ctx, d_chunk_off);
    bs_stat_add(ctx, "sign_propagation", prop ? 1.0 : 0.0);
    bs_stat_add(ctx, "sign_brute_force", brute ? 1.0 : 0.0);
    bs_stat_add(ctx, "n_active", (double)h_sc[0]);
    bs_stat_add(ctx, "n_sign_seeds", (double)(prop ? h_sc[1] : h_sc[0]));
    bs_stat_add(ctx, "n_sign_items", (double)n_items);
    bs_stat_add(ctx, "n_heavy_bricks", (double)n_heavy);
    bs_stat_add(ctx, "n_heavy_items", (double)n_hitems);
    BS_CUDA(ctx, cudaGetLastError());
    return BS_OK;
}
