// Flood fill + CSG (union / subtract / intersect) on sorted bricks (sm_100a data path, host-side directory).
// Replaces Volume::{union, intersect, subtract} (src/voxel/volume/mod.rs:74-93):
//   FloodFill  leaf_node/flood_fill.rs:12-70, internal_node/flood_fill.rs:17-117, root_node/flood_fill.rs:8-63
//   Csg        leaf_node/csg.rs:17-45, internal_node/csg.rs:20-163, root_node/csg.rs:9-58
//
// The reference runs both over its pointer tree. Here the tree's upper levels (root map -> 32^3 node -> 16^3
// node) are re-derived from the sorted brick keys as a small host-side directory: per node one byte per slot
// (kind: inactive / child / active tile, and the sign the reference's flood fill would leave there, including
// the quirk that an internal node's last_value_sign() is the FIRST value sign of its last child). The CSG rules
// of internal_node/csg.rs are evaluated slot by slot on that directory and produce a sorted list of output
// bricks, each tagged copy-A / copy-B / copy-B-negated / merge. The data path is one kernel: a CTA per output
// brick re-derives the leaf flood fill of its operands on the fly (inactive voxels = +-f32::MAX with the
// scan-line sign of leaf_node/flood_fill.rs) and applies min / max(a,-b) / max over all 512 slots, mask |= mask.
// Per merged brick: read 2 x 2112 B, write 2112 B.
//
// Not reproduced (documented in DESIGN.md): a 16^3 node that a
// subtract/intersect emptied stays in the reference's tree with stale background signs read by later flood
// fills (dangling union bytes, undefined in the reference) -> here it disappears.
#include "bs_common.cuh"
#include <algorithm>
#include <cfloat>
#include <cstring>

namespace {

typedef unsigned long long u64;
typedef unsigned char u8;

enum { K_INACTIVE = 0, K_CHILD = 1, K_TILE = 2 };
inline u8 mk(int kind, bool neg) { return (u8)((kind << 1) | (neg ? 1 : 0)); }
inline int kind_of(u8 s) { return s >> 1; }
inline bool neg_of(u8 s) { return s & 1; }
inline bool inside_tile(u8 s) { return kind_of(s) != K_CHILD && neg_of(s); }    // internal_node/csg.rs:20-23
inline bool outside_tile(u8 s) { return kind_of(s) != K_CHILD && !neg_of(s); }  // :25-28

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

// ---- host directory --------------------------------------------------------------------------------------------
struct Dir {
    std::vector<u64> bkeys; std::vector<u8> bfirst, blast;
    std::vector<u64> t8k, t128k; std::vector<float> t8v, t128v;
    std::vector<u64> n4k; std::vector<u8> n4state, n4first, n4last;     // per 16^3 node: 4096 slot bytes
    std::vector<u64> n5k; std::vector<u8> n5state, n5first, n5last;     // per 32^3 node: 32768 slot bytes
    long find(const std::vector<u64>& v, u64 k) const { auto it = std::lower_bound(v.begin(), v.end(), k); return (it != v.end() && *it == k) ? (long)(it - v.begin()) : -1; }
    u8 s5(long n5, unsigned slot) const { return n5state[(size_t)n5 * 32768 + slot]; }
    u8 s4(long n4, unsigned slot) const { return n4state[(size_t)n4 * 4096 + slot]; }
};

// Scan-line flood fill over the R^3 slots of an internal node (internal_node/flood_fill.rs:17-75). `state` holds
// kind for occupied slots (and tile signs); child_first/child_last give a child's first/last value sign.
template <class FirstFn, class LastFn>
bool flood_internal(u8* state, int log2, FirstFn child_first, LastFn child_last) {
    const int R = 1 << log2, SIZE = R * R * R;
    int fo = -1;
    for (int o = 0; o < SIZE; ++o) if (kind_of(state[o]) != K_INACTIVE) { fo = o; break; }
    if (fo < 0) return false;
    // (Some(v), Some(b)) if v <= b => tile sign, else first branch's first_value_sign (:31-39): the lowest occupied slot decides
    bool i = kind_of(state[fo]) == K_TILE ? neg_of(state[fo]) : child_first(fo);
    auto running = [&](int o, bool cur) { const int k = kind_of(state[o]); return k == K_CHILD ? child_last(o) : (k == K_TILE ? neg_of(state[o]) : cur); };
    for (int x = 0; x < R; ++x) {
        const int x00 = x << (2 * log2);
        i = running(x00, i);
        bool j = i;
        for (int y = 0; y < R; ++y) {
            const int xy0 = x00 + (y << log2);
            j = running(xy0, j);
            bool k = j;
            for (int z = 0; z < R; ++z) {
                const int o = xy0 + z;
                if (kind_of(state[o]) == K_INACTIVE) state[o] = mk(K_INACTIVE, k); else k = running(o, k);
            }
        }
    }
    return true;
}

bs_status build_dir(bs_context* ctx, const bs_volume* v, Dir& D) {
    cudaStream_t st = ctx->stream;
    const size_t n = v->n_bricks;
    D.bkeys.resize(n); D.bfirst.resize(n); D.blast.resize(n);
    D.t8k.resize(v->n_tiles8); D.t8v.resize(v->n_tiles8); D.t128k.resize(v->n_tiles128); D.t128v.resize(v->n_tiles128);
    u8 *d_first = nullptr, *d_last = nullptr;
    BS_TRY(bs_alloc(ctx, &d_first, n)); BS_TRY(bs_alloc(ctx, &d_last, n));
    if (n) {
        bs_count_launch(), k_brick_signs<<<(unsigned)n, 512, 0, st>>>(v->values, v->masks, d_first, d_last);
        BS_CUDA(ctx, cudaMemcpyAsync(D.bkeys.data(), v->keys, n * sizeof(u64), cudaMemcpyDeviceToHost, st));
        BS_CUDA(ctx, cudaMemcpyAsync(D.bfirst.data(), d_first, n, cudaMemcpyDeviceToHost, st));
        BS_CUDA(ctx, cudaMemcpyAsync(D.blast.data(), d_last, n, cudaMemcpyDeviceToHost, st));
    }
    if (v->n_tiles8) { BS_CUDA(ctx, cudaMemcpyAsync(D.t8k.data(), v->tile8_keys, v->n_tiles8 * sizeof(u64), cudaMemcpyDeviceToHost, st)); BS_CUDA(ctx, cudaMemcpyAsync(D.t8v.data(), v->tile8_values, v->n_tiles8 * sizeof(float), cudaMemcpyDeviceToHost, st)); }
    if (v->n_tiles128) { BS_CUDA(ctx, cudaMemcpyAsync(D.t128k.data(), v->tile128_keys, v->n_tiles128 * sizeof(u64), cudaMemcpyDeviceToHost, st)); BS_CUDA(ctx, cudaMemcpyAsync(D.t128v.data(), v->tile128_values, v->n_tiles128 * sizeof(float), cudaMemcpyDeviceToHost, st)); }
    BS_CUDA(ctx, cudaStreamSynchronize(st));
    bs_free(ctx, d_first); bs_free(ctx, d_last);
    // 16^3 nodes
    for (u64 k : D.bkeys) if (D.n4k.empty() || D.n4k.back() != (k >> 12)) D.n4k.push_back(k >> 12);
    if (!D.t8k.empty()) {
        std::vector<u64> t; for (u64 k : D.t8k) t.push_back(k >> 12);
        std::vector<u64> m; std::merge(D.n4k.begin(), D.n4k.end(), t.begin(), t.end(), std::back_inserter(m));
        m.erase(std::unique(m.begin(), m.end()), m.end()); D.n4k.swap(m);
    }
    const size_t n4 = D.n4k.size();
    D.n4state.assign(n4 * 4096, mk(K_INACTIVE, false)); D.n4first.assign(n4, 0); D.n4last.assign(n4, 0);
    std::vector<int> slot_brick(4096);
    size_t bi = 0, ti = 0;
    for (size_t j = 0; j < n4; ++j) {
        u8* s = &D.n4state[j * 4096];
        const size_t b0 = bi;
        for (; bi < n && (D.bkeys[bi] >> 12) == D.n4k[j]; ++bi) { s[D.bkeys[bi] & 4095] = mk(K_CHILD, false); slot_brick[D.bkeys[bi] & 4095] = (int)bi; }
        for (; ti < D.t8k.size() && (D.t8k[ti] >> 12) == D.n4k[j]; ++ti) s[D.t8k[ti] & 4095] = mk(K_TILE, std::signbit(D.t8v[ti]));
        (void)b0;
        flood_internal(s, 4, [&](int o) { return D.bfirst[slot_brick[o]] != 0; }, [&](int o) { return D.blast[slot_brick[o]] != 0; });
        D.n4first[j] = kind_of(s[0]) == K_CHILD ? D.bfirst[slot_brick[0]] : (u8)neg_of(s[0]);
        D.n4last[j] = kind_of(s[4095]) == K_CHILD ? D.bfirst[slot_brick[4095]] : (u8)neg_of(s[4095]);  // quirk: FIRST sign of the last child (:103-108)
    }
    // 32^3 nodes
    for (u64 k : D.n4k) if (D.n5k.empty() || D.n5k.back() != (k >> 15)) D.n5k.push_back(k >> 15);
    if (!D.t128k.empty()) {
        std::vector<u64> t; for (u64 k : D.t128k) t.push_back(k >> 15);
        std::vector<u64> m; std::merge(D.n5k.begin(), D.n5k.end(), t.begin(), t.end(), std::back_inserter(m));
        m.erase(std::unique(m.begin(), m.end()), m.end()); D.n5k.swap(m);
    }
    const size_t n5 = D.n5k.size();
    D.n5state.assign(n5 * 32768, mk(K_INACTIVE, false)); D.n5first.assign(n5, 0); D.n5last.assign(n5, 0);
    std::vector<int> slot_n4(32768);
    size_t ci = 0; ti = 0;
    for (size_t j = 0; j < n5; ++j) {
        u8* s = &D.n5state[j * 32768];
        for (; ci < n4 && (D.n4k[ci] >> 15) == D.n5k[j]; ++ci) { s[D.n4k[ci] & 32767] = mk(K_CHILD, false); slot_n4[D.n4k[ci] & 32767] = (int)ci; }
        for (; ti < D.t128k.size() && (D.t128k[ti] >> 15) == D.n5k[j]; ++ti) s[D.t128k[ti] & 32767] = mk(K_TILE, std::signbit(D.t128v[ti]));
        flood_internal(s, 5, [&](int o) { return D.n4first[slot_n4[o]] != 0; }, [&](int o) { return D.n4last[slot_n4[o]] != 0; });
        D.n5first[j] = kind_of(s[0]) == K_CHILD ? D.n4first[slot_n4[0]] : (u8)neg_of(s[0]);
        D.n5last[j] = kind_of(s[32767]) == K_CHILD ? D.n4first[slot_n4[32767]] : (u8)neg_of(s[32767]);
    }
    // root flood fill (root_node/flood_fill.rs:17-41): two consecutive 4096^3 nodes on the same z-line that are not
    // adjacent and face each other with negative signs get every key between them filled with an empty node whose
    // background is negative: all of its slots then read as inside tiles in the CSG rules below
    std::vector<u64> ins;
    for (size_t j = 0; j + 1 < n5; ++j) {
        const u64 a = D.n5k[j], b = D.n5k[j + 1];
        if ((a >> 9) != (b >> 9) || (b & 511) == (a & 511) + 1) continue;
        if (!(D.n5last[j] && D.n5first[j + 1])) continue;
        for (u64 z = (a & 511) + 1; z < (b & 511); ++z) ins.push_back((a & ~511ull) | z);
    }
    if (!ins.empty()) {
        if (ins.size() > 4096) return bs_fail(ctx, BS_ERR_UNSUPPORTED, "root-level flood fill would insert %zu empty 4096^3 nodes", ins.size());
        std::vector<u64> k2; std::vector<u8> st2, f2, l2;
        k2.reserve(n5 + ins.size()); st2.reserve((n5 + ins.size()) * 32768);
        size_t ia = 0, ib = 0;
        while (ia < n5 || ib < ins.size()) {
            if (ib >= ins.size() || (ia < n5 && D.n5k[ia] < ins[ib])) {
                k2.push_back(D.n5k[ia]); st2.insert(st2.end(), D.n5state.begin() + ia * 32768, D.n5state.begin() + (ia + 1) * 32768);
                f2.push_back(D.n5first[ia]); l2.push_back(D.n5last[ia]); ++ia;
            } else {
                k2.push_back(ins[ib]); st2.insert(st2.end(), 32768, mk(K_INACTIVE, true)); f2.push_back(1); l2.push_back(1); ++ib;
            }
        }
        D.n5k.swap(k2); D.n5state.swap(st2); D.n5first.swap(f2); D.n5last.swap(l2);
    }
    return BS_OK;
}

struct OutTile { u64 key; float value; };

}  // namespace

bs_status bs_csg_impl(bs_volume* A, bs_volume* B, int op, bs_volume** out) {
    bs_context* ctx = A->ctx;
    cudaStream_t st = ctx->stream;
    bs_marks_begin(ctx);
    Dir a, b;
    BS_TRY(build_dir(ctx, A, a));
    BS_TRY(build_dir(ctx, B, b));
    bs_mark(ctx, "csg_flood_fill_ms");
    enum { UNION = 0, SUBTRACT = 1, INTERSECT = 2 };
    const int merge_mode = 3 + op;
    std::vector<OutBrick> ra, rb;
    std::vector<OutTile> t8, t128;
    // --- bricks and tiles of A -------------------------------------------------------------------------------
    // fate of something of A living under 32^3 slot `slot5` (A has a CHILD there): 0 drop, 1 keep, 2 recurse
    auto a_level5 = [&](long bn5, unsigned slot5) -> int {
        if (bn5 < 0) return op == INTERSECT ? 0 : 1;                       // root_node/csg.rs: key only in self
        const u8 sb = b.s5(bn5, slot5);
        if (op == UNION) return inside_tile(sb) ? 0 : (kind_of(sb) == K_CHILD ? 2 : 1);
        if (op == SUBTRACT) return outside_tile(sb) ? 1 : (inside_tile(sb) ? 0 : 2);
        return inside_tile(sb) ? 1 : (outside_tile(sb) ? 0 : 2);
    };
    for (size_t i = 0; i < a.bkeys.size(); ++i) {
        const u64 k = a.bkeys[i];
        const long bn5 = b.find(b.n5k, k >> 27);
        int f = a_level5(bn5, (unsigned)((k >> 12) & 32767));
        if (f == 2) {
            const long bn4 = b.find(b.n4k, k >> 12);
            const u8 sb = b.s4(bn4, (unsigned)(k & 4095));
            if (kind_of(sb) == K_CHILD) { ra.push_back({k, (int)i, (int)b.find(b.bkeys, k), merge_mode}); continue; }
            if (op == UNION) f = inside_tile(sb) ? 0 : 1;
            else if (op == SUBTRACT) f = outside_tile(sb) ? 1 : 0;
            else f = inside_tile(sb) ? 1 : 0;
        }
        if (f == 1) ra.push_back({k, (int)i, -1, 0});
    }
    for (size_t i = 0; i < a.t8k.size(); ++i) {  // active 8^3 tiles of A (A's 16^3 slot is a tile)
        const u64 k = a.t8k[i];
        const bool an = std::signbit(a.t8v[i]);
        const long bn5 = b.find(b.n5k, k >> 27);
        int f = a_level5(bn5, (unsigned)((k >> 12) & 32767));
        if (f == 2) {
            const u8 sb = b.s4(b.find(b.n4k, k >> 12), (unsigned)(k & 4095));
            if (op == UNION) f = an ? 1 : ((inside_tile(sb) || kind_of(sb) == K_CHILD) ? 0 : 1);
            else if (op == SUBTRACT) f = !an ? 1 : (outside_tile(sb) ? 1 : 0);
            else f = !an ? 1 : (inside_tile(sb) ? 1 : 0);
        }
        if (f == 1) t8.push_back({k, a.t8v[i]});
    }
    for (size_t i = 0; i < a.t128k.size(); ++i) {  // active 128^3 tiles of A
        const u64 k = a.t128k[i];
        const bool an = std::signbit(a.t128v[i]);
        const long bn5 = b.find(b.n5k, k >> 15);
        int f;
        if (bn5 < 0) f = op == INTERSECT ? 0 : 1;
        else {
            const u8 sb = b.s5(bn5, (unsigned)(k & 32767));
            if (op == UNION) f = an ? 1 : ((inside_tile(sb) || kind_of(sb) == K_CHILD) ? 0 : 1);
            else if (op == SUBTRACT) f = !an ? 1 : (outside_tile(sb) ? 1 : 0);
            else f = !an ? 1 : (inside_tile(sb) ? 1 : 0);
        }
        if (f == 1) t128.push_back({k, a.t128v[i]});
    }
    // --- bricks and tiles of B -------------------------------------------------------------------------------
    // fate of something of B under 32^3 slot `slot5` (B has a CHILD there): 0 drop, 1 take, 2 take negated, 3 recurse
    auto b_level5 = [&](long an5, unsigned slot5) -> int {
        if (an5 < 0) return op == UNION ? 1 : 0;                             // key only in other
        const u8 sa = a.s5(an5, slot5);
        if (op == UNION) return inside_tile(sa) ? 0 : (outside_tile(sa) ? 1 : 3);
        if (op == SUBTRACT) return outside_tile(sa) ? 0 : (inside_tile(sa) ? 2 : 3);
        return outside_tile(sa) ? 0 : (inside_tile(sa) ? 1 : 3);
    };
    for (size_t i = 0; i < b.bkeys.size(); ++i) {
        const u64 k = b.bkeys[i];
        int f = b_level5(a.find(a.n5k, k >> 27), (unsigned)((k >> 12) & 32767));
        if (f == 3) {
            const u8 sa = a.s4(a.find(a.n4k, k >> 12), (unsigned)(k & 4095));
            if (kind_of(sa) == K_CHILD) continue;  // merged brick, emitted from A's side
            if (op == UNION) f = inside_tile(sa) ? 0 : 1;
            else if (op == SUBTRACT) f = outside_tile(sa) ? 0 : 2;
            else f = outside_tile(sa) ? 0 : 1;
        }
        if (f == 1) rb.push_back({k, -1, (int)i, 1});
        else if (f == 2) rb.push_back({k, -1, (int)i, 2});
    }
    for (size_t i = 0; i < b.t8k.size(); ++i) {  // B's 8^3 tiles travel only with a whole 16^3 node taken at the 32^3 level
        const u64 k = b.t8k[i];
        const int f = b_level5(a.find(a.n5k, k >> 27), (unsigned)((k >> 12) & 32767));
        if (f == 1) t8.push_back({k, b.t8v[i]});
        else if (f == 2) t8.push_back({k, -b.t8v[i]});
    }
    for (size_t i = 0; i < b.t128k.size(); ++i)  // B's 128^3 tiles travel only with a whole 32^3 node (union, key only in other)
        if (op == UNION && a.find(a.n5k, b.t128k[i] >> 15) < 0) t128.push_back({b.t128k[i], b.t128v[i]});
    // --- union: make_child_inside (internal_node/csg.rs:42-48,74-77) creates ACTIVE -MAX tiles ----------------
    if (op == UNION) {
        for (size_t ja = 0; ja < a.n5k.size(); ++ja) {
            const long jb = b.find(b.n5k, a.n5k[ja]);
            if (jb < 0) continue;
            for (unsigned s5 = 0; s5 < 32768; ++s5) {
                const u8 sa = a.s5((long)ja, s5), sb = b.s5(jb, s5);
                if (inside_tile(sa)) continue;
                if (inside_tile(sb)) { t128.push_back({(a.n5k[ja] << 15) | s5, -FLT_MAX}); continue; }
                if (kind_of(sa) == K_CHILD && kind_of(sb) == K_CHILD) {
                    const u64 k4 = (a.n5k[ja] << 15) | s5;
                    const long a4 = a.find(a.n4k, k4), b4 = b.find(b.n4k, k4);
                    for (unsigned s4 = 0; s4 < 4096; ++s4)
                        if (!inside_tile(a.s4(a4, s4)) && inside_tile(b.s4(b4, s4))) t8.push_back({(k4 << 12) | s4, -FLT_MAX});
                }
            }
        }
    }
    std::vector<OutBrick> recs(ra.size() + rb.size());
    std::merge(ra.begin(), ra.end(), rb.begin(), rb.end(), recs.begin(), [](const OutBrick& x, const OutBrick& y) { return x.key < y.key; });
    std::sort(t8.begin(), t8.end(), [](const OutTile& x, const OutTile& y) { return x.key < y.key; });
    std::sort(t128.begin(), t128.end(), [](const OutTile& x, const OutTile& y) { return x.key < y.key; });
    bs_mark(ctx, "csg_directory_ms");
    // --- data path ------------------------------------------------------------------------------------------------
    bs_volume* R = bs_volume_new(ctx, A->voxel_size);
    bs_status s = bs_volume_alloc_bricks(R, recs.size());
    OutBrick* d_recs = nullptr;
    if (s == BS_OK) s = bs_alloc(ctx, &d_recs, recs.size());
    if (s == BS_OK && !recs.empty()) {
        cudaMemcpyAsync(d_recs, recs.data(), recs.size() * sizeof(OutBrick), cudaMemcpyHostToDevice, st);
        bs_count_launch(), k_csg_bricks<<<(unsigned)recs.size(), 512, 0, st>>>(d_recs, A->values, A->masks, B->values, B->masks, R->keys, R->values, R->masks);
    }
    auto upload_tiles = [&](const std::vector<OutTile>& t, size_t& n, u64*& keys, float*& vals) -> bs_status {
        n = t.size();
        if (!n) return BS_OK;
        std::vector<u64> k(n); std::vector<float> v(n);
        for (size_t i = 0; i < n; ++i) { k[i] = t[i].key; v[i] = t[i].value; }
        BS_TRY(bs_alloc(ctx, &keys, n)); BS_TRY(bs_alloc(ctx, &vals, n));
        BS_CUDA(ctx, cudaMemcpyAsync(keys, k.data(), n * sizeof(u64), cudaMemcpyHostToDevice, st));
        BS_CUDA(ctx, cudaMemcpyAsync(vals, v.data(), n * sizeof(float), cudaMemcpyHostToDevice, st));
        BS_CUDA(ctx, cudaStreamSynchronize(st));  // the staging vectors die at scope exit
        return BS_OK;
    };
    if (s == BS_OK) s = upload_tiles(t8, R->n_tiles8, R->tile8_keys, R->tile8_values);
    if (s == BS_OK) s = upload_tiles(t128, R->n_tiles128, R->tile128_keys, R->tile128_values);
    bs_mark(ctx, "csg_bricks_ms");
    if (s == BS_OK && (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess)) s = bs_fail(ctx, BS_ERR_CUDA, "csg kernels failed");
    bs_free(ctx, d_recs);
    if (s != BS_OK) { bs_volume_free(R); return s; }
    bs_marks_end(ctx);
    size_t n_merge = 0; for (auto& r : recs) if (r.mode >= 3) ++n_merge;
    bs_stat_add(ctx, "n_out_bricks", (double)recs.size());
    bs_stat_add(ctx, "n_merged_bricks", (double)n_merge);
    bs_stat_add(ctx, "n_out_tiles", (double)(t8.size() + t128.size()));
    *out = R;
    return BS_OK;
}
