// Flood fill + CSG (union / subtract / intersect) on sorted bricks (sm_100a; data path AND directory on the device).
// Replaces Volume::{union, intersect, subtract} (src/voxel/volume/mod.rs:74-93):
//   FloodFill  leaf_node/flood_fill.rs:12-70, internal_node/flood_fill.rs:17-117, root_node/flood_fill.rs:8-63
//   Csg        leaf_node/csg.rs:17-45, internal_node/csg.rs:20-163, root_node/csg.rs:9-58
//
// The reference runs both over its pointer tree. Here the tree's upper levels (root map -> 32^3 node -> 16^3
// node) are re-derived from the sorted brick keys as a small directory in HBM: per node one byte per slot
// (kind: inactive / child / active tile, and the sign the reference's flood fill would leave there, including
// the quirk that an internal node's last_value_sign() is the FIRST value sign of its last child), filled by one CTA
// per node (k_dir_flood). The CSG rules of internal_node/csg.rs are evaluated by one thread per brick / tile / slot
// on the two directories and produce a sorted list of output bricks, each tagged copy-A / copy-B / copy-B-negated /
// merge. Nodes that the root-level flood fill inserts between two inside nodes are kept as key ranges (no limit). The data path is one kernel: a CTA per output
// brick re-derives the leaf flood fill of its operands on the fly (inactive voxels = +-f32::MAX with the
// scan-line sign of leaf_node/flood_fill.rs) and applies min / max(a,-b) / max over all 512 slots, mask |= mask.
// Per merged brick: read 2 x 2112 B, write 2112 B.
//
// Not reproduced (documented in DESIGN.md): a 16^3 node that a
// subtract/intersect emptied stays in the reference's tree with stale background signs read by later flood
// fills (dangling union bytes, undefined in the reference) -> here it disappears.
#include "bs_common.cuh"
#include <cub/cub.cuh>
#include <algorithm>
#include <cfloat>
#include <cstring>

namespace {

typedef unsigned long long u64;
typedef unsigned char u8;

enum { K_INACTIVE = 0, K_CHILD = 1, K_TILE = 2 };

// ---- device: leaf flood fill signs ----------------------------------------------------------------------------
// Sign (1 = negative) the reference's leaf flood fill gives voxel t of a brick (512 threads, one per voxel).
// leaf_node/flood_fill.rs:12-50: inactive (x,y,z) takes the sign of the last active voxel before it on its z-line,
// else of the last active (x,y',0), y' <= y, else of the last active (x',0,0), x' <= x, else of the first active
// voxel of the leaf. Active voxels keep their own sign.
__device__ unsigned brick_fill_sign(const float v, const bool act, const u64* __restrict__ mask8, const float* __restrict__ vals,
                                    unsigned* s_a0 /*8: per x, bit y = active(x,y,0)*/, unsigned* s_n0 /*8: sign bits*/, unsigned* s_first) {
    const unsigned t = threadIdx.x, lane = t & 31, x = t >> 6, y = (t >> 3) & 7, z = t & 7;
    const unsigned neg = __float_as_uint(v) >> 31;
    const unsigned ba = __ballot_sync(0xFFFFFFFFu, act), bn = __ballot_sync(0xFFFFFFFFu, neg != 0);
    if (t < 8) { s_a0[t] = 0; s_n0[t] = 0; }
    if (t == 0) {
        unsigned first = 0;  // sign of the first active voxel (lowest offset); 0 if the brick is empty
        for (int w = 0; w < 8; ++w) { const u64 m = mask8[w]; if (m) { const int b = __ffsll((long long)m) - 1; first = __float_as_uint(vals[w * 64 + b]) >> 31; break; } }
        *s_first = first;
    }
    __syncthreads();
    if (z == 0 && act) { atomicOr(&s_a0[x], 1u << y); if (neg) atomicOr(&s_n0[x], 1u << y); }
    __syncthreads();
    if (act) return neg;
    const unsigned line = lane & ~7u;
    const unsigned la = (ba >> line) & ((2u << z) - 1u) & 0xFFu;
    if (la) { const unsigned zz = 31 - __clz(la); return (bn >> (line + zz)) & 1u; }
    const unsigned ya = s_a0[x] & ((2u << y) - 1u);
    if (ya) { const unsigned yy = 31 - __clz(ya); return (s_n0[x] >> yy) & 1u; }
    for (int xx = (int)x; xx >= 0; --xx) if (s_a0[xx] & 1u) return s_n0[xx] & 1u;
    return *s_first;
}

// first_value_sign / last_value_sign of every brick after its flood fill (leaf_node/flood_fill.rs:60-68)
__global__ void __launch_bounds__(512) k_brick_signs(const float* __restrict__ values, const u64* __restrict__ masks, u8* first, u8* last) {
    __shared__ unsigned s_a0[8], s_n0[8], s_first;
    const size_t b = blockIdx.x;
    const unsigned t = threadIdx.x;
    const float v = values[b * 512 + t];
    const bool act = (masks[b * 8 + (t >> 6)] >> (t & 63)) & 1;
    const unsigned s = brick_fill_sign(v, act, masks + b * 8, values + b * 512, s_a0, s_n0, &s_first);
    if (t == 0) first[b] = (u8)s;
    if (t == 511) last[b] = (u8)s;
}

struct OutBrick { u64 key; int a, b, mode; };  // mode 0 copy A, 1 copy B, 2 copy -B, 3 union, 4 subtract, 5 intersect

__global__ void __launch_bounds__(512) k_csg_bricks(const OutBrick* __restrict__ recs, const float* __restrict__ va, const u64* __restrict__ ma,
                                                    const float* __restrict__ vb, const u64* __restrict__ mb, u64* keys, float* values, u64* masks) {
    __shared__ unsigned s_a0[8], s_n0[8], s_first;
    __shared__ unsigned s_bal[16];
    const size_t o = blockIdx.x;
    const OutBrick r = recs[o];
    const unsigned t = threadIdx.x;
    if (t == 0) keys[o] = r.key;
    float out; bool act;
    if (r.mode <= 2) {
        const float* sv = r.mode == 0 ? va + (size_t)r.a * 512 : vb + (size_t)r.b * 512;
        const u64* sm = r.mode == 0 ? ma + (size_t)r.a * 8 : mb + (size_t)r.b * 8;
        out = sv[t]; act = (sm[t >> 6] >> (t & 63)) & 1;
        if (r.mode == 2) out = -out;  // Csg::flip_signs (leaf_node/csg.rs:41-45)
    } else {
        const float a = va[(size_t)r.a * 512 + t], b = vb[(size_t)r.b * 512 + t];
        const bool aa = (ma[(size_t)r.a * 8 + (t >> 6)] >> (t & 63)) & 1, ab = (mb[(size_t)r.b * 8 + (t >> 6)] >> (t & 63)) & 1;
        const unsigned sa = brick_fill_sign(a, aa, ma + (size_t)r.a * 8, va + (size_t)r.a * 512, s_a0, s_n0, &s_first);
        __syncthreads();
        const unsigned sb = brick_fill_sign(b, ab, mb + (size_t)r.b * 8, vb + (size_t)r.b * 512, s_a0, s_n0, &s_first);
        const float fa = aa ? a : (sa ? -FLT_MAX : FLT_MAX), fb = ab ? b : (sb ? -FLT_MAX : FLT_MAX);
        // partial_min / partial_max (voxel/utils.rs:39-61): the second operand wins unless strictly less / greater
        if (r.mode == 3) out = (fa < fb) ? fa : fb;
        else if (r.mode == 4) { const float nb = -fb; out = (fa > nb) ? fa : nb; }
        else out = (fa > fb) ? fa : fb;
        act = aa || ab;
    }
    values[o * 512 + t] = out;
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, act);
    if ((t & 31) == 0) s_bal[t >> 5] = bal;
    __syncthreads();
    if (t < 8) masks[o * 8 + t] = (u64)s_bal[2 * t] | ((u64)s_bal[2 * t + 1] << 32);
}

// ---- device directory -------------------------------------------------------------------------------------------
// Slot byte: bit 0 = sign (1 = negative), bits 1-2 = kind, bit 3 / 4 = first / last value sign of the child in a CHILD slot.
__host__ __device__ __forceinline__ u8 mk(int kind, bool neg) { return (u8)((kind << 1) | (neg ? 1 : 0)); }
__host__ __device__ __forceinline__ int kind_of(u8 s) { return (s >> 1) & 3; }
__host__ __device__ __forceinline__ bool neg_of(u8 s) { return s & 1; }
__host__ __device__ __forceinline__ bool inside_tile(u8 s) { return kind_of(s) != K_CHILD && neg_of(s); }    // internal_node/csg.rs:20-23
__host__ __device__ __forceinline__ bool outside_tile(u8 s) { return kind_of(s) != K_CHILD && !neg_of(s); }  // :25-28
__device__ __forceinline__ u8 mk_child(bool first, bool last) { return (u8)((K_CHILD << 1) | (first ? 8 : 0) | (last ? 16 : 0)); }
__device__ __forceinline__ bool child_first(u8 s) { return (s >> 3) & 1; }
__device__ __forceinline__ bool child_last(u8 s) { return (s >> 4) & 1; }

struct Dir {  // one operand: everything lives in HBM
    size_t n = 0, nt8 = 0, nt128 = 0, n4 = 0, n5 = 0;
    const u64* bkeys = nullptr; const u64 *t8k = nullptr, *t128k = nullptr; const float *t8v = nullptr, *t128v = nullptr;
    u8 *bfirst = nullptr, *blast = nullptr;
    u64 *n4k = nullptr, *n5k = nullptr;                   // sorted keys of the 16^3 / 32^3 nodes
    u8 *s4 = nullptr, *s5 = nullptr;                      // slot bytes: [n4][4096], [n5][32768]
    u8 *n4first = nullptr, *n4last = nullptr, *n5first = nullptr, *n5last = nullptr;
    // root flood fill (root_node/flood_fill.rs:17-41): gap j = the keys strictly between n5k[j] and n5k[j+1] on one z-line
    // that the reference fills with empty all-negative nodes; kept as ranges [gap_lo[j], gap_hi[j]) (empty when lo >= hi)
    u64 *gap_lo = nullptr, *gap_hi = nullptr;
};
__device__ __forceinline__ long dev_find(const u64* __restrict__ v, size_t n, u64 k) {
    size_t lo = 0, hi = n;
    while (lo < hi) { const size_t mid = (lo + hi) >> 1; if (v[mid] < k) lo = mid + 1; else hi = mid; }
    return (lo < n && v[lo] == k) ? (long)lo : -1;
}
// 32^3 node with key k5: index >= 0, -2 = node inserted by the root flood fill (every slot an inside tile), -1 = none
__device__ __forceinline__ long dir_find5(const Dir& D, u64 k5) {
    size_t lo = 0, hi = D.n5;  // first index with key > k5
    while (lo < hi) { const size_t mid = (lo + hi) >> 1; if (D.n5k[mid] <= k5) lo = mid + 1; else hi = mid; }
    if (lo == 0) return -1;
    const size_t j = lo - 1;
    if (D.n5k[j] == k5) return (long)j;
    if (j + 1 < D.n5 && D.gap_lo[j] <= k5 && k5 < D.gap_hi[j]) return -2;
    return -1;
}
__device__ __forceinline__ u8 dir_s5(const Dir& D, long n5, unsigned slot) { return n5 == -2 ? mk(K_INACTIVE, true) : D.s5[(size_t)n5 * 32768 + slot]; }
__device__ __forceinline__ u8 dir_s4(const Dir& D, long n4, unsigned slot) { return D.s4[(size_t)n4 * 4096 + slot]; }

__global__ void k_dir_node_keys(const u64* __restrict__ a, size_t na, const u64* __restrict__ b, size_t nb, int shift, u64* out) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < na) out[i] = a[i] >> shift; else if (i < na + nb) out[i] = b[i - na] >> shift;
}
__global__ void k_dir_scatter4(Dir D) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < D.n) { const u64 k = D.bkeys[i]; D.s4[(size_t)dev_find(D.n4k, D.n4, k >> 12) * 4096 + (k & 4095)] = mk_child(D.bfirst[i], D.blast[i]); }
    else if (i < D.n + D.nt8) { const u64 k = D.t8k[i - D.n]; D.s4[(size_t)dev_find(D.n4k, D.n4, k >> 12) * 4096 + (k & 4095)] = mk(K_TILE, __float_as_uint(D.t8v[i - D.n]) >> 31); }
}
__global__ void k_dir_scatter5(Dir D) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < D.n4) { const u64 k = D.n4k[i]; D.s5[(size_t)dev_find(D.n5k, D.n5, k >> 15) * 32768 + (k & 32767)] = mk_child(D.n4first[i], D.n4last[i]); }
    else if (i < D.n4 + D.nt128) { const u64 k = D.t128k[i - D.n4]; D.s5[(size_t)dev_find(D.n5k, D.n5, k >> 15) * 32768 + (k & 32767)] = mk(K_TILE, __float_as_uint(D.t128v[i - D.n4]) >> 31); }
}
// Scan-line flood fill over the R^3 slots of one internal node (internal_node/flood_fill.rs:17-75), one CTA per node, one
// thread per (x, y) line: an inactive slot (x, y, z) takes the sign of the last occupied slot before it on its z-line, else
// of the last occupied (x, y', 0), y' <= y, else of the last occupied (x', 0, 0), x' <= x, else of the node's first occupied
// slot (a tile's sign or a child's FIRST value sign, :31-39); an occupied slot hands on a tile's sign or a child's LAST value
// sign. first / last value sign of the node (:85-108): slot 0 / slot R^3 - 1 -- with the reference's quirk that
// last_value_sign() of an internal node is the FIRST value sign of its last child.
template <int L> __global__ void __launch_bounds__(1 << (2 * L)) k_dir_flood(u8* state, u8* nfirst, u8* nlast) {
    constexpr int R = 1 << L, SIZE = R * R * R;
    __shared__ unsigned s_occ[R], s_sgn[R], s_fo;
    u8* s = state + (size_t)blockIdx.x * SIZE;
    const unsigned t = threadIdx.x, x = t >> L, y = t & (R - 1);
    if (t < R) { s_occ[t] = 0; s_sgn[t] = 0; }
    if (t == 0) s_fo = 0xFFFFFFFFu;
    __syncthreads();
    u8 line[R];
    {   // R consecutive bytes
        const uint4* p = reinterpret_cast<const uint4*>(s + (size_t)t * R);
#pragma unroll
        for (int q = 0; q < R / 16; ++q) { const uint4 v = p[q]; memcpy(line + 16 * q, &v, 16); }
    }
    auto hand_on = [](u8 b) -> unsigned { return kind_of(b) == K_CHILD ? (unsigned)child_last(b) : (unsigned)neg_of(b); };
    int fz = -1;
#pragma unroll
    for (int z = R - 1; z >= 0; --z) if (kind_of(line[z]) != K_INACTIVE) fz = z;
    if (fz >= 0) atomicMin(&s_fo, t * R + (unsigned)fz);
    if (kind_of(line[0]) != K_INACTIVE) { atomicOr(&s_occ[x], 1u << y); if (hand_on(line[0])) atomicOr(&s_sgn[x], 1u << y); }
    __syncthreads();
    if (s_fo == 0xFFFFFFFFu) return;  // empty node (cannot happen: a node exists because something lives in it)
    unsigned k;
    {
        const u8 f = s[s_fo];
        unsigned i = kind_of(f) == K_TILE ? (unsigned)neg_of(f) : (unsigned)child_first(f);
        for (int xx = (int)x; xx >= 0; --xx) if (s_occ[xx] & 1u) { i = s_sgn[xx] & 1u; break; }
        const unsigned ya = s_occ[x] & ((2u << y) - 1u);  // (2u << 31 == 0: all bits)
        k = ya ? (s_sgn[x] >> (31 - __clz(ya))) & 1u : i;
    }
#pragma unroll
    for (int z = 0; z < R; ++z) { if (kind_of(line[z]) == K_INACTIVE) line[z] = mk(K_INACTIVE, k); else k = hand_on(line[z]); }
    {
        uint4* p = reinterpret_cast<uint4*>(s + (size_t)t * R);
#pragma unroll
        for (int q = 0; q < R / 16; ++q) { uint4 v; memcpy(&v, line + 16 * q, 16); p[q] = v; }
    }
    if (t == 0) nfirst[blockIdx.x] = kind_of(line[0]) == K_CHILD ? (u8)child_first(line[0]) : (u8)neg_of(line[0]);
    if (t == R * R - 1) nlast[blockIdx.x] = kind_of(line[R - 1]) == K_CHILD ? (u8)child_first(line[R - 1]) : (u8)neg_of(line[R - 1]);
}
__global__ void k_dir_gaps(Dir D) {
    const size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (j >= D.n5) return;
    u64 lo = 1, hi = 0;
    if (j + 1 < D.n5) {
        const u64 a = D.n5k[j], b = D.n5k[j + 1];
        if ((a >> 9) == (b >> 9) && (b & 511) != (a & 511) + 1 && D.n5last[j] && D.n5first[j + 1]) { lo = a + 1; hi = b; }
    }
    D.gap_lo[j] = lo; D.gap_hi[j] = hi;
}

bs_status sort_unique(bs_context* ctx, u64* d_in, size_t m, int bits, u64** d_out, size_t* n_out) {
    cudaStream_t st = ctx->stream;
    u64 *d_sorted = nullptr, *d_uniq = nullptr; size_t* d_n = nullptr; void* d_tmp = nullptr; size_t tmp = 0, tmp2 = 0;
    BS_TRY(bs_alloc(ctx, &d_sorted, m)); BS_TRY(bs_alloc(ctx, &d_uniq, m)); BS_TRY(bs_alloc(ctx, &d_n, 1));
    cub::DeviceRadixSort::SortKeys(nullptr, tmp, d_in, d_sorted, m, 0, bits, st);
    cub::DeviceSelect::Unique(nullptr, tmp2, d_sorted, d_uniq, d_n, m, st);
    BS_TRY(bs_alloc(ctx, (char**)&d_tmp, std::max(tmp, tmp2)));
    cub::DeviceRadixSort::SortKeys(d_tmp, tmp, d_in, d_sorted, m, 0, bits, st);
    cub::DeviceSelect::Unique(d_tmp, tmp2, d_sorted, d_uniq, d_n, m, st);
    BS_CUDA(ctx, cudaMemcpyAsync(n_out, d_n, sizeof(size_t), cudaMemcpyDeviceToHost, st));
    BS_CUDA(ctx, cudaStreamSynchronize(st));
    bs_free(ctx, d_tmp); bs_free(ctx, d_sorted); bs_free(ctx, d_n);
    *d_out = d_uniq;
    return BS_OK;
}

bs_status build_dir(bs_context* ctx, const bs_volume* v, Dir& D) {
    cudaStream_t st = ctx->stream;
    D.n = v->n_bricks; D.nt8 = v->n_tiles8; D.nt128 = v->n_tiles128;
    D.bkeys = v->keys; D.t8k = v->tile8_keys; D.t8v = v->tile8_values; D.t128k = v->tile128_keys; D.t128v = v->tile128_values;
    BS_TRY(bs_alloc(ctx, &D.bfirst, D.n)); BS_TRY(bs_alloc(ctx, &D.blast, D.n));
    if (D.n) bs_count_launch(), k_brick_signs<<<(unsigned)D.n, 512, 0, st>>>(v->values, v->masks, D.bfirst, D.blast);
    // 16^3 nodes: everything that holds a brick or an 8^3 tile
    if (D.n + D.nt8) {
        u64* d_k = nullptr;
        BS_TRY(bs_alloc(ctx, &d_k, D.n + D.nt8));
        bs_count_launch(), k_dir_node_keys<<<bs_blocks(D.n + D.nt8, 256), 256, 0, st>>>(D.bkeys, D.n, D.t8k, D.nt8, 12, d_k);
        BS_TRY(sort_unique(ctx, d_k, D.n + D.nt8, 42, &D.n4k, &D.n4));
        bs_free(ctx, d_k);
        BS_TRY(bs_alloc(ctx, &D.s4, D.n4 * 4096)); BS_TRY(bs_alloc(ctx, &D.n4first, D.n4)); BS_TRY(bs_alloc(ctx, &D.n4last, D.n4));
        BS_CUDA(ctx, cudaMemsetAsync(D.s4, 0, D.n4 * 4096, st));
        bs_count_launch(), k_dir_scatter4<<<bs_blocks(D.n + D.nt8, 256), 256, 0, st>>>(D);
        bs_count_launch(), k_dir_flood<4><<<(unsigned)D.n4, 256, 0, st>>>(D.s4, D.n4first, D.n4last);
    }
    // 32^3 nodes: everything that holds a 16^3 node or a 128^3 tile
    if (D.n4 + D.nt128) {
        u64* d_k = nullptr;
        BS_TRY(bs_alloc(ctx, &d_k, D.n4 + D.nt128));
        bs_count_launch(), k_dir_node_keys<<<bs_blocks(D.n4 + D.nt128, 256), 256, 0, st>>>(D.n4k, D.n4, D.t128k, D.nt128, 15, d_k);
        BS_TRY(sort_unique(ctx, d_k, D.n4 + D.nt128, 27, &D.n5k, &D.n5));
        bs_free(ctx, d_k);
        BS_TRY(bs_alloc(ctx, &D.s5, D.n5 * 32768)); BS_TRY(bs_alloc(ctx, &D.n5first, D.n5)); BS_TRY(bs_alloc(ctx, &D.n5last, D.n5));
        BS_TRY(bs_alloc(ctx, &D.gap_lo, D.n5)); BS_TRY(bs_alloc(ctx, &D.gap_hi, D.n5));
        BS_CUDA(ctx, cudaMemsetAsync(D.s5, 0, D.n5 * 32768, st));
        bs_count_launch(), k_dir_scatter5<<<bs_blocks(D.n4 + D.nt128, 256), 256, 0, st>>>(D);
        bs_count_launch(), k_dir_flood<5><<<(unsigned)D.n5, 1024, 0, st>>>(D.s5, D.n5first, D.n5last);
        bs_count_launch(), k_dir_gaps<<<bs_blocks(D.n5, 256), 256, 0, st>>>(D);
    }
    BS_CUDA(ctx, cudaGetLastError());
    return BS_OK;
}
void free_dir(bs_context* ctx, Dir& D) {
    bs_free(ctx, D.bfirst); bs_free(ctx, D.blast); bs_free(ctx, D.n4k); bs_free(ctx, D.n5k); bs_free(ctx, D.s4); bs_free(ctx, D.s5);
    bs_free(ctx, D.n4first); bs_free(ctx, D.n4last); bs_free(ctx, D.n5first); bs_free(ctx, D.n5last); bs_free(ctx, D.gap_lo); bs_free(ctx, D.gap_hi);
    D = Dir();
}

// ---- CSG rules on the two directories (internal_node/csg.rs:68-152, root_node/csg.rs:9-58) --------------------------------
enum { UNION = 0, SUBTRACT = 1, INTERSECT = 2 };
struct OutTile { u64 key; float value; };
// fate of something of A living under 32^3 slot `slot5` (A has a CHILD there): 0 drop, 1 keep, 2 recurse
__device__ __forceinline__ int a_level5(const Dir& B, int op, long bn5, unsigned slot5) {
    if (bn5 == -1) return op == INTERSECT ? 0 : 1;  // root_node/csg.rs: key only in self
    const u8 sb = dir_s5(B, bn5, slot5);
    if (op == UNION) return inside_tile(sb) ? 0 : (kind_of(sb) == K_CHILD ? 2 : 1);
    if (op == SUBTRACT) return outside_tile(sb) ? 1 : (inside_tile(sb) ? 0 : 2);
    return inside_tile(sb) ? 1 : (outside_tile(sb) ? 0 : 2);
}
// fate of something of B under 32^3 slot `slot5` (B has a CHILD there): 0 drop, 1 take, 2 take negated, 3 recurse
__device__ __forceinline__ int b_level5(const Dir& A, int op, long an5, unsigned slot5) {
    if (an5 == -1) return op == UNION ? 1 : 0;  // key only in other
    const u8 sa = dir_s5(A, an5, slot5);
    if (op == UNION) return inside_tile(sa) ? 0 : (outside_tile(sa) ? 1 : 3);
    if (op == SUBTRACT) return outside_tile(sa) ? 0 : (inside_tile(sa) ? 2 : 3);
    return outside_tile(sa) ? 0 : (inside_tile(sa) ? 1 : 3);
}
// one thread per brick of A, then of B: output record (or key = invalid)
__global__ void k_csg_brick_fates(Dir A, Dir B, int op, OutBrick* recs) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= A.n + B.n) return;
    OutBrick r{BS_KEY_INVALID, -1, -1, 0};
    if (i < A.n) {
        const u64 k = A.bkeys[i];
        int f = a_level5(B, op, dir_find5(B, k >> 27), (unsigned)((k >> 12) & 32767));
        if (f == 2) {
            const u8 sb = dir_s4(B, dev_find(B.n4k, B.n4, k >> 12), (unsigned)(k & 4095));
            if (kind_of(sb) == K_CHILD) { r = OutBrick{k, (int)i, (int)dev_find(B.bkeys, B.n, k), 3 + op}; f = -1; }
            else if (op == UNION) f = inside_tile(sb) ? 0 : 1;
            else if (op == SUBTRACT) f = outside_tile(sb) ? 1 : 0;
            else f = inside_tile(sb) ? 1 : 0;
        }
        if (f == 1) r = OutBrick{k, (int)i, -1, 0};
    } else {
        const size_t ib = i - A.n;
        const u64 k = B.bkeys[ib];
        int f = b_level5(A, op, dir_find5(A, k >> 27), (unsigned)((k >> 12) & 32767));
        if (f == 3) {
            const u8 sa = dir_s4(A, dev_find(A.n4k, A.n4, k >> 12), (unsigned)(k & 4095));
            if (kind_of(sa) == K_CHILD) f = 0;  // merged brick, emitted from A's side
            else if (op == UNION) f = inside_tile(sa) ? 0 : 1;
            else if (op == SUBTRACT) f = outside_tile(sa) ? 0 : 2;
            else f = outside_tile(sa) ? 0 : 1;
        }
        if (f == 1) r = OutBrick{k, -1, (int)ib, 1};
        else if (f == 2) r = OutBrick{k, -1, (int)ib, 2};
    }
    recs[i] = r;
}
__device__ __forceinline__ void tile_push(OutTile* out, unsigned* count, unsigned cap, u64 key, float value) {
    const unsigned slot = atomicAdd(count, 1u);
    if (slot < cap) out[slot] = OutTile{key, value};
}
// one thread per active tile of A (8^3 then 128^3), then of B; kept tiles are appended (sorted afterwards)
__global__ void k_csg_tile_fates(Dir A, Dir B, int op, OutTile* t8, unsigned* n_t8, unsigned cap8, OutTile* t128, unsigned* n_t128, unsigned cap128) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < A.nt8) {  // A's 16^3 slot is a tile
        const u64 k = A.t8k[i]; const float v = A.t8v[i];
        const bool an = __float_as_uint(v) >> 31;
        int f = a_level5(B, op, dir_find5(B, k >> 27), (unsigned)((k >> 12) & 32767));
        if (f == 2) {
            const u8 sb = dir_s4(B, dev_find(B.n4k, B.n4, k >> 12), (unsigned)(k & 4095));
            if (op == UNION) f = an ? 1 : ((inside_tile(sb) || kind_of(sb) == K_CHILD) ? 0 : 1);
            else if (op == SUBTRACT) f = !an ? 1 : (outside_tile(sb) ? 1 : 0);
            else f = !an ? 1 : (inside_tile(sb) ? 1 : 0);
        }
        if (f == 1) tile_push(t8, n_t8, cap8, k, v);
        return;
    }
    i -= A.nt8;
    if (i < A.nt128) {
        const u64 k = A.t128k[i]; const float v = A.t128v[i];
        const bool an = __float_as_uint(v) >> 31;
        const long bn5 = dir_find5(B, k >> 15);
        int f;
        if (bn5 == -1) f = op == INTERSECT ? 0 : 1;
        else {
            const u8 sb = dir_s5(B, bn5, (unsigned)(k & 32767));
            if (op == UNION) f = an ? 1 : ((inside_tile(sb) || kind_of(sb) == K_CHILD) ? 0 : 1);
            else if (op == SUBTRACT) f = !an ? 1 : (outside_tile(sb) ? 1 : 0);
            else f = !an ? 1 : (inside_tile(sb) ? 1 : 0);
        }
        if (f == 1) tile_push(t128, n_t128, cap128, k, v);
        return;
    }
    i -= A.nt128;
    if (i < B.nt8) {  // B's 8^3 tiles travel only with a whole 16^3 node taken at the 32^3 level
        const u64 k = B.t8k[i];
        const int f = b_level5(A, op, dir_find5(A, k >> 27), (unsigned)((k >> 12) & 32767));
        if (f == 1) tile_push(t8, n_t8, cap8, k, B.t8v[i]);
        else if (f == 2) tile_push(t8, n_t8, cap8, k, -B.t8v[i]);
        return;
    }
    i -= B.nt8;
    if (i < B.nt128) {  // B's 128^3 tiles travel only with a whole 32^3 node (union, key only in other)
        if (op == UNION && dir_find5(A, B.t128k[i] >> 15) == -1) tile_push(t128, n_t128, cap128, B.t128k[i], B.t128v[i]);
    }
}
// union: make_child_inside (internal_node/csg.rs:42-48,74-77) creates ACTIVE -MAX tiles. LEVEL 5: one thread per slot of
// every 32^3 node of A that B also has (a real node or one inserted by its root flood fill); LEVEL 4: per slot of every
// 16^3 node of A whose counterpart exists in B (both sides hold a child in the 32^3 slot).
template <int LEVEL> __global__ void k_csg_union_tiles(Dir A, Dir B, OutTile* out, unsigned* count, unsigned cap) {
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (LEVEL == 5) {
        const size_t ja = g >> 15; const unsigned s5 = (unsigned)(g & 32767);
        if (ja >= A.n5) return;
        const long jb = dir_find5(B, A.n5k[ja]);
        if (jb == -1) return;
        const u8 sa = A.s5[ja * 32768 + s5], sb = dir_s5(B, jb, s5);
        if (!inside_tile(sa) && inside_tile(sb)) tile_push(out, count, cap, (A.n5k[ja] << 15) | s5, -FLT_MAX);
    } else {
        const size_t ja = g >> 12; const unsigned s4 = (unsigned)(g & 4095);
        if (ja >= A.n4) return;
        const long jb = dev_find(B.n4k, B.n4, A.n4k[ja]);
        if (jb < 0) return;
        if (!inside_tile(A.s4[ja * 4096 + s4]) && inside_tile(B.s4[(size_t)jb * 4096 + s4])) tile_push(out, count, cap, (A.n4k[ja] << 12) | s4, -FLT_MAX);
    }
}
struct RecValid { __device__ bool operator()(const OutBrick& r) const { return r.key != BS_KEY_INVALID; } };
__global__ void k_rec_keys(const OutBrick* __restrict__ recs, size_t n, u64* keys, unsigned* idx) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = recs[i].key; idx[i] = (unsigned)i; }
}
__global__ void k_rec_gather(const OutBrick* __restrict__ recs, const unsigned* __restrict__ order, size_t n, OutBrick* out, unsigned* n_merge) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const OutBrick r = recs[order[i]];
    out[i] = r;
    if (r.mode >= 3) atomicAdd(n_merge, 1u);
}
__global__ void k_tile_split(const OutTile* __restrict__ t, size_t n, u64* keys, float* vals) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = t[i].key; vals[i] = t[i].value; }
}

// sorts n tiles by key into freshly allocated key / value arrays of the result volume
bs_status finish_tiles(bs_context* ctx, const OutTile* d_t, size_t n, int bits, u64** keys, float** vals) {
    cudaStream_t st = ctx->stream;
    *keys = nullptr; *vals = nullptr;
    if (!n) return BS_OK;
    u64* d_k = nullptr; float* d_v = nullptr; void* d_tmp = nullptr; size_t tmp = 0;
    BS_TRY(bs_alloc(ctx, &d_k, n)); BS_TRY(bs_alloc(ctx, &d_v, n)); BS_TRY(bs_alloc(ctx, keys, n)); BS_TRY(bs_alloc(ctx, vals, n));
    bs_count_launch(), k_tile_split<<<bs_blocks(n, 256), 256, 0, st>>>(d_t, n, d_k, d_v);
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, d_k, *keys, d_v, *vals, n, 0, bits, st);
    BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp));
    cub::DeviceRadixSort::SortPairs(d_tmp, tmp, d_k, *keys, d_v, *vals, n, 0, bits, st);
    bs_free(ctx, d_tmp); bs_free(ctx, d_k); bs_free(ctx, d_v);
    return BS_OK;
}

}  // namespace

bs_status bs_csg_impl(bs_volume* A, bs_volume* B, int op, bs_volume** out) {
    bs_context* ctx = A->ctx;
    cudaStream_t st = ctx->stream;
    bs_marks_begin(ctx);
    Dir a, b;
    BS_TRY(build_dir(ctx, A, a));
    BS_TRY(build_dir(ctx, B, b));
    bs_mark(ctx, "csg_flood_fill_ms");
    // --- fates of all bricks; tiles kept / created; everything stays on the device ------------------------------------
    const size_t n_in = a.n + b.n;
    OutBrick *d_fate = nullptr, *d_kept = nullptr, *d_recs = nullptr; size_t* d_nkept = nullptr; unsigned* d_cnt = nullptr;  // d_cnt: [0] 8^3 tiles, [1] 128^3 tiles, [2] merged bricks
    BS_TRY(bs_alloc(ctx, &d_fate, n_in)); BS_TRY(bs_alloc(ctx, &d_kept, n_in)); BS_TRY(bs_alloc(ctx, &d_nkept, 1)); BS_TRY(bs_alloc(ctx, &d_cnt, 4));
    BS_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, 4 * sizeof(unsigned), st));
    if (n_in) bs_count_launch(), k_csg_brick_fates<<<bs_blocks(n_in, 256), 256, 0, st>>>(a, b, op, d_fate);
    void* d_tmp = nullptr; size_t tmp = 0;
    cub::DeviceSelect::If(nullptr, tmp, d_fate, d_kept, d_nkept, n_in, RecValid(), st);
    BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp));
    cub::DeviceSelect::If(d_tmp, tmp, d_fate, d_kept, d_nkept, n_in, RecValid(), st);
    bs_free(ctx, d_tmp); d_tmp = nullptr;
    // tiles: every kept tile of A / B, plus (union) one per slot that make_child_inside turns into an inside tile
    const size_t n_tile_in = a.nt8 + a.nt128 + b.nt8 + b.nt128;
    unsigned cap8 = (unsigned)(a.nt8 + b.nt8), cap128 = (unsigned)(a.nt128 + b.nt128);
    OutTile *d_t8 = nullptr, *d_t128 = nullptr;
    unsigned h_cnt[4] = {0, 0, 0, 0};
    if (op == UNION) {  // count the created tiles first (the buffers are sized exactly)
        if (a.n5) bs_count_launch(), k_csg_union_tiles<5><<<bs_blocks(a.n5 * 32768, 256), 256, 0, st>>>(a, b, nullptr, d_cnt + 1, 0u);
        if (a.n4) bs_count_launch(), k_csg_union_tiles<4><<<bs_blocks(a.n4 * 4096, 256), 256, 0, st>>>(a, b, nullptr, d_cnt, 0u);
        BS_CUDA(ctx, cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
    }
    size_t n_kept = 0;
    BS_CUDA(ctx, cudaMemcpyAsync(&n_kept, d_nkept, sizeof(size_t), cudaMemcpyDeviceToHost, st));
    BS_CUDA(ctx, cudaStreamSynchronize(st));
    cap8 += h_cnt[0]; cap128 += h_cnt[1];
    BS_TRY(bs_alloc(ctx, &d_t8, (size_t)cap8)); BS_TRY(bs_alloc(ctx, &d_t128, (size_t)cap128));
    BS_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, 4 * sizeof(unsigned), st));
    if (n_tile_in) bs_count_launch(), k_csg_tile_fates<<<bs_blocks(n_tile_in, 256), 256, 0, st>>>(a, b, op, d_t8, d_cnt, cap8, d_t128, d_cnt + 1, cap128);
    if (op == UNION && h_cnt[1]) bs_count_launch(), k_csg_union_tiles<5><<<bs_blocks(a.n5 * 32768, 256), 256, 0, st>>>(a, b, d_t128, d_cnt + 1, cap128);
    if (op == UNION && h_cnt[0]) bs_count_launch(), k_csg_union_tiles<4><<<bs_blocks(a.n4 * 4096, 256), 256, 0, st>>>(a, b, d_t8, d_cnt, cap8);
    // output bricks in key order (A's and B's survivors are each ascending; a radix sort of the concatenation merges them)
    bs_volume* R = bs_volume_new(ctx, A->voxel_size);
    bs_status s = bs_volume_alloc_bricks(R, n_kept);
    if (s != BS_OK) { bs_volume_free(R); return s; }
    if (n_kept) {
        u64 *d_k = nullptr, *d_k2 = nullptr; unsigned *d_i = nullptr, *d_i2 = nullptr;
        BS_TRY(bs_alloc(ctx, &d_k, n_kept)); BS_TRY(bs_alloc(ctx, &d_k2, n_kept)); BS_TRY(bs_alloc(ctx, &d_i, n_kept)); BS_TRY(bs_alloc(ctx, &d_i2, n_kept)); BS_TRY(bs_alloc(ctx, &d_recs, n_kept));
        bs_count_launch(), k_rec_keys<<<bs_blocks(n_kept, 256), 256, 0, st>>>(d_kept, n_kept, d_k, d_i);
        tmp = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp, d_k, d_k2, d_i, d_i2, n_kept, 0, 54, st);
        BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp));
        cub::DeviceRadixSort::SortPairs(d_tmp, tmp, d_k, d_k2, d_i, d_i2, n_kept, 0, 54, st);
        bs_count_launch(), k_rec_gather<<<bs_blocks(n_kept, 256), 256, 0, st>>>(d_kept, d_i2, n_kept, d_recs, d_cnt + 2);
        bs_free(ctx, d_tmp); bs_free(ctx, d_k); bs_free(ctx, d_k2); bs_free(ctx, d_i); bs_free(ctx, d_i2);
    }
    BS_CUDA(ctx, cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
    bs_mark(ctx, "csg_directory_ms");
    // --- data path ------------------------------------------------------------------------------------------------
    if (n_kept) bs_count_launch(), k_csg_bricks<<<(unsigned)n_kept, 512, 0, st>>>(d_recs, A->values, A->masks, B->values, B->masks, R->keys, R->values, R->masks);
    bs_mark(ctx, "csg_bricks_ms");
    if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) { bs_volume_free(R); return bs_fail(ctx, BS_ERR_CUDA, "csg kernels failed"); }
    R->n_tiles8 = h_cnt[0]; R->n_tiles128 = h_cnt[1];
    s = finish_tiles(ctx, d_t8, R->n_tiles8, 54, &R->tile8_keys, &R->tile8_values);
    if (s == BS_OK) s = finish_tiles(ctx, d_t128, R->n_tiles128, 42, &R->tile128_keys, &R->tile128_values);
    if (s == BS_OK && cudaStreamSynchronize(st) != cudaSuccess) s = bs_fail(ctx, BS_ERR_CUDA, "csg tile sort failed");
    bs_free(ctx, d_fate); bs_free(ctx, d_kept); bs_free(ctx, d_recs); bs_free(ctx, d_nkept); bs_free(ctx, d_cnt); bs_free(ctx, d_t8); bs_free(ctx, d_t128);
    free_dir(ctx, a); free_dir(ctx, b);
    if (s != BS_OK) { bs_volume_free(R); return s; }
    bs_marks_end(ctx);
    bs_stat_add(ctx, "n_out_bricks", (double)n_kept);
    bs_stat_add(ctx, "n_merged_bricks", (double)h_cnt[2]);
    bs_stat_add(ctx, "n_out_tiles", (double)(R->n_tiles8 + R->n_tiles128));
    *out = R;
    return BS_OK;
}
