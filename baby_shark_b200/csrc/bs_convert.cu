// Mesh -> unsigned narrow-band distance field on sorted 8^3 bricks (sm_100a).
// Replaces MeshToVolume::{subdivide_triangle, compute_unsigned_distance_field}
// (src/voxel/mesh_to_volume.rs:75-196) and Triangle3::{bbox, closest_point, max_side}
// (src/geometry/primitives/triangle3.rs:117-124,307-382).
//
// value(v) = min over sub-triangles WHOSE INTEGER BOX CONTAINS v of |closest_point(v) - v|   (not the true
// distance: mesh_to_volume.rs:170-195), topology = union of those boxes. Sub-triangle vertices come from
// sequential f32 running sums in the reference (:90-115); every thread re-derives its sub-triangle with the
// same additions in the same order, so boxes and distances are bit-identical and the scatter-min (strict
// `<`, order independent) can be an atomicMin on the float bits (distances are >= 0).
//
// Two passes over the (never materialised) sub-triangle stream:
//   mark : insert the key of every brick a box touches into an open-addressing hash set
//   eval : point-triangle distances for every lattice point of every box, atomicMin into the brick
// with a radix sort of the brick keys in between (sorted key order == the reference's leaf visit order).
#include "bs_common.cuh"
#include "bs_ptdist.cuh"
#include <cub/cub.cuh>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstring>

namespace {

constexpr int TPB = 256;
constexpr int MAX_PROBE = 512;

struct ConvertParams {
    const float* tris; size_t n_tris;   // n_tris = triangles this call works on (entries of tri_ids when that is set)
    const unsigned* tri_ids;            // sharded runs: the triangles that can touch this rank's bricks (null: all, in order)
    const unsigned long long* offsets;  // exclusive prefix sum of per-triangle sub-triangle counts, n_tris + 1
    const unsigned* cta_start;          // triangle holding sub-triangle b * TPB, per CTA b (k_cta_starts): saves every CTA of k_mark / k_eval a 23-step dependent search
    unsigned* dummy_brick;              // 512 scratch words: where lattice points of bricks this rank does not keep are sent (sharded runs)
    unsigned long long total;
    float vs, inv_vs; int band;
    unsigned long long* table_keys; unsigned* table_slots; unsigned table_mask;
    unsigned* table_counts;  // per hash slot: how many sub-triangle boxes touch the brick (load-balancing weight for sharding); may be null
    float* values;
    int* flags;  // [0] hash overflow, [1] index range error
    unsigned* derr;  // bs_context::d_err
    unsigned long long* n_eval;  // sum of box volumes = point-triangle evaluations (roofline work counter)
    int use_clip, clip_mn[3], clip_mx[3];  // sharded runs: voxel bounding box of the bricks this rank keeps
};

__device__ __forceinline__ const float* tri_ptr(const ConvertParams& P, size_t t) { return P.tris + 9 * (size_t)(P.tri_ids ? P.tri_ids[t] : t); }
__device__ __forceinline__ unsigned long long hash64(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k;
}

__device__ __forceinline__ f3 ld3(const float* p) { return {p[0], p[1], p[2]}; }


// Triangle3::max_side / voxel_size, floored (mesh_to_volume.rs:76)
__device__ __forceinline__ float num_subs_of(f3 p1, f3 p2, f3 p3, float vs) {
    float ab = xnorm2(xsub(p2, p1)), ac = xnorm2(xsub(p3, p1)), bc = xnorm2(xsub(p3, p2));
    float m = fmaxf(fmaxf(ab, ac), bc);
    return floorf(xdiv(xsqrt(m), vs));
}

__global__ void k_tri_counts(const float* __restrict__ tris, const unsigned* __restrict__ tri_ids, size_t n_tris, float vs, unsigned long long* counts, double* area_vox) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    double a = 0.0;
    if (t < n_tris) {
        const float* p = tris + 9 * (size_t)(tri_ids ? tri_ids[t] : t);
        f3 p1 = ld3(p), p2 = ld3(p + 3), p3 = ld3(p + 6);
        float n = num_subs_of(p1, p2, p3, vs);
        unsigned long long c;
        if (n < 2.0f) c = 1; else if (n != n) c = 0; else { double nd = fmin((double)n, 4.0e9); c = (unsigned long long)(nd * nd); }
        counts[t] = c;
        f3 cr = xcross(xsub(p2, p1), xsub(p3, p1));
        float ar = 0.5f * sqrtf(xnorm2(cr));
        if (ar == ar && ar < 3.0e38f) a = (double)ar / ((double)vs * (double)vs);
    }
    // block reduce -> one atomic per block
    typedef cub::BlockReduce<double, TPB> BR;
    __shared__ typename BR::TempStorage tmp;
    double s = BR(tmp).Sum(a);
    if (threadIdx.x == 0 && s != 0.0) atomicAdd(area_vox, s);
}

// j-th sub-triangle of triangle (p1,p2,p3) in the reference's construction (mesh_to_volume.rs:75-116):
// row i = isqrt(j); inside the row, pos = j - i*i: pos == 2i is the row's closing triangle (a,b,c), otherwise
// step k = pos >> 1 yields (a_prev, b_s, a_s) for even pos and (a_s, b_s, c_s) for odd pos.
__device__ void make_subtri(f3 p1, f3 p2, f3 p3, float vs, unsigned long long j, f3& A, f3& B, f3& C) {
    float n = num_subs_of(p1, p2, p3, vs);
    if (n < 2.0f) { A = p1; B = p2; C = p3; return; }
    float inv = xdiv(1.0f, n);
    f3 s1 = xscale(xsub(p2, p1), inv), s2 = xscale(xsub(p3, p2), inv);
    unsigned long long i = (unsigned long long)sqrt((double)j);
    while (i * i > j) --i;
    while ((i + 1) * (i + 1) <= j) ++i;
    unsigned long long pos = j - i * i;
    f3 a = p1;
    for (unsigned long long r = 0; r < i; ++r) a = xadd(a, s1);
    f3 b = xadd(a, s1);
    if (pos == 2 * i) { A = a; B = b; C = xadd(b, s2); return; }
    unsigned long long k = pos >> 1;
    f3 a_s = xadd(a, s2), b_s = xadd(b, s2), a_prev = a;
    for (unsigned long long q = 0; q < k; ++q) { a_prev = a_s; a_s = xadd(a_s, s2); b_s = xadd(b_s, s2); }
    if ((pos & 1) == 0) { A = a_prev; B = b_s; C = a_s; }
    else { A = a_s; B = b_s; C = xadd(b_s, s2); }  // c_s(k) = b_s(k) + s2 bit for bit (c = b + s2)
}

// Integer box of a sub-triangle (mesh_to_volume.rs:124-141). Returns false when outside the supported range.
__device__ __forceinline__ bool subtri_box(f3 A, f3 B, f3 C, float inv_vs, int band, int mn[3], int mx[3]) {
    float lo[3] = {fminf(C.x, fminf(A.x, B.x)), fminf(C.y, fminf(A.y, B.y)), fminf(C.z, fminf(A.z, B.z))};
    float hi[3] = {fmaxf(C.x, fmaxf(A.x, B.x)), fmaxf(C.y, fmaxf(A.y, B.y)), fmaxf(C.z, fmaxf(A.z, B.z))};
    bool ok = true;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float fl = floorf(xmul(lo[d], inv_vs)), ce = ceilf(xmul(hi[d], inv_vs));
        if (!(fl > -1.0e6f && ce < 1.0e6f)) { ok = false; fl = 0.f; ce = 0.f; }
        mn[d] = (int)fl - band; mx[d] = (int)ce + band;
    }
    if (mx[0] == mn[0] || mx[1] == mn[1] || mx[2] == mn[2]) {
#pragma unroll
        for (int d = 0; d < 3; ++d) { mn[d] -= 1; mx[d] += 1; }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) if ((mn[d] >> 3) < BS_BRICK_MIN || (mx[d] >> 3) > BS_BRICK_MAX) ok = false;
    return ok;
}

// Block-cooperative map from a flat sub-triangle index to (triangle, local index): "last t with
// offsets[t] <= g" (zero-count triangles repeat an offset and are skipped by taking the LAST such t).
struct TriCursor { size_t tri; unsigned long long local; bool valid; };
__device__ __forceinline__ size_t find_tri(const ConvertParams& P, unsigned long long g) {
    size_t lo = 0, hi = P.n_tris;  // offsets[0] = 0 <= g < total = offsets[n_tris]
    while (hi - lo > 1) { size_t mid = (lo + hi) >> 1; if (P.offsets[mid] <= g) lo = mid; else hi = mid; }
    return lo;
}
__device__ TriCursor locate(const ConvertParams& P, unsigned long long* s_off /*[TPB+1]*/) {
    const unsigned long long g0 = (unsigned long long)blockIdx.x * TPB;
    // uniform search for the triangle containing g0 (every thread walks the same path: broadcast loads), then
    // a window of TPB+1 offsets in shared memory serves the per-thread searches
    const size_t base = P.cta_start ? (size_t)P.cta_start[blockIdx.x] : find_tri(P, g0);
    for (int i = threadIdx.x; i <= TPB; i += TPB) { size_t idx = base + i; s_off[i] = idx <= P.n_tris ? P.offsets[idx] : ~0ull; }
    __syncthreads();
    TriCursor c; c.valid = false; c.tri = 0; c.local = 0;
    unsigned long long g = g0 + threadIdx.x;
    if (g < P.total) {
        int l = 0, h = TPB + 1;
        while (h - l > 1) { int m = (l + h) >> 1; if (s_off[m] <= g) l = m; else h = m; }
        size_t t = (l == TPB) ? find_tri(P, g) : base + l;  // window exhausted only if zero-count triangles intervene
        c.tri = t; c.local = g - P.offsets[t]; c.valid = true;
    }
    return c;
}

__global__ void k_cta_starts(ConvertParams P, unsigned n_ctas, unsigned* start) {
    const unsigned b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n_ctas) start[b] = (unsigned)find_tri(P, (unsigned long long)b * TPB);
}

__device__ __forceinline__ void hash_insert(const ConvertParams& P, unsigned long long key, unsigned count) {
    unsigned h = (unsigned)hash64(key) & P.table_mask;
    for (int probe = 0; probe < MAX_PROBE; ++probe) {
        unsigned long long cur = P.table_keys[h];
        bool mine = cur == key;
        if (!mine && cur == BS_KEY_INVALID) {
            unsigned long long prev = atomicCAS(&P.table_keys[h], BS_KEY_INVALID, key);
            mine = prev == BS_KEY_INVALID || prev == key;
        }
        if (mine) { if (P.table_counts) atomicAdd(&P.table_counts[h], count); return; }
        h = (h + 1) & P.table_mask;
    }
    P.flags[0] = 1;
}
__device__ __forceinline__ unsigned hash_lookup(const ConvertParams& P, unsigned long long key) {
    unsigned h = (unsigned)hash64(key) & P.table_mask;
    for (int probe = 0; probe < MAX_PROBE; ++probe) {
        unsigned long long cur = P.table_keys[h];
        if (cur == key) return P.table_slots[h];
        if (cur == BS_KEY_INVALID) return 0xFFFFFFFFu;
        h = (h + 1) & P.table_mask;
    }
    atomicOr(P.derr, BS_DERR_PROBE);  // cannot happen after a mark pass without overflow (every insert ended within MAX_PROBE); never a silent miss
    return 0xFFFFFFFFu;
}

__global__ void __launch_bounds__(TPB) k_mark(ConvertParams P) {
    __shared__ unsigned long long s_off[TPB + 1];
    TriCursor c = locate(P, s_off);
    unsigned long long vol = 0;
    int mn[3] = {0, 0, 0}, mx[3] = {-1, -1, -1};
    if (c.valid) {
        const float* p = tri_ptr(P, c.tri);
        f3 A, B, C;
        make_subtri(ld3(p), ld3(p + 3), ld3(p + 6), P.vs, c.local, A, B, C);
        if (!subtri_box(A, B, C, P.inv_vs, P.band, mn, mx)) { P.flags[1] = 1; mx[0] = mn[0] - 1; }
        else vol = (unsigned long long)(mx[0] - mn[0] + 1) * (unsigned long long)(mx[1] - mn[1] + 1) * (unsigned long long)(mx[2] - mn[2] + 1);
    }
    for (int o = 16; o; o >>= 1) vol += __shfl_xor_sync(0xFFFFFFFFu, vol, o);
    if ((threadIdx.x & 31) == 0 && vol) atomicAdd(P.n_eval, vol);
    const bool has = c.valid && mx[0] >= mn[0];
    const bool wide = has && ((mx[0] >> 3) - (mn[0] >> 3) > 1 || (mx[1] >> 3) - (mn[1] >> 3) > 1 || (mx[2] >> 3) - (mn[2] >> 3) > 1);
    if (__any_sync(0xFFFFFFFFu, wide)) {  // large narrow bands: a box spans more than 2 bricks along an axis
        if (!has) return;
        for (int bx = mn[0] >> 3; bx <= (mx[0] >> 3); ++bx)
            for (int by = mn[1] >> 3; by <= (mx[1] >> 3); ++by)
                for (int bz = mn[2] >> 3; bz <= (mx[2] >> 3); ++bz) hash_insert(P, bs_brick_key(bx, by, bz), 1u);
        return;
    }
    // <= 2 x 2 x 2 bricks per box. The 32 sub-triangles of a warp are neighbours on the surface and mostly touch the same few bricks:
    // lanes holding the same key elect one of them, which inserts the key once and adds the whole group to the brick's touch count
    const int bx0 = mn[0] >> 3, by0 = mn[1] >> 3, bz0 = mn[2] >> 3;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int bx = bx0 + (k & 1), by = by0 + ((k >> 1) & 1), bz = bz0 + (k >> 2);
        const bool touched = has && bx <= (mx[0] >> 3) && by <= (mx[1] >> 3) && bz <= (mx[2] >> 3);
        if (!__any_sync(0xFFFFFFFFu, touched)) continue;
        const unsigned long long key = touched ? bs_brick_key(bx, by, bz) : BS_KEY_INVALID;
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, key);
        if (touched && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) hash_insert(P, key, (unsigned)__popc(peers));
    }
}

// Distances. Each lane derives one sub-triangle (vertices, box, the <= 2x2x2 brick slots its box touches) and parks it in
// shared memory; then the warp flattens the (sub-triangle, z-column) pairs of its 32 sub-triangles and deals them out
// round robin, so lanes stay busy although box sizes differ (a plain thread-per-sub-triangle loop ran at 10 of 32 lanes:
// r1 ncu). The point-triangle distance is bs_ptdist.cuh: the x / y halves of the six dot products are computed once per
// column, the seven regions are predicates around one shared division instead of a chain of divergent branches (the
// branchy form ran at 20 of 32 lanes), records are read with 128-bit shared loads, the column's (x, y) comes from a
// multiply-shift instead of a division, its two possible brick slots (a box spans at most two bricks along z) are fetched
// once, and the scatter-min works on SQUARED distances: sqrt is monotone, so k_masks (bs_fwn.cu), which reads every value
// anyway, takes the root of the minimum. Boxes wider than 9 voxels (large narrow bands) take the per-thread path.
// Measured and rejected (config 5, 6.6 ms): reading the current value first and skipping the RED when it is not lower
// (7.3 ms through L2 or L1: the kernel is issue-bound, 603 M REDs cost nothing extra -- without any RED: 6.5 ms);
// 3 CTAs / SM at 78 registers (7.4 ms).
constexpr int EV2_REC = 28;  // words per parked sub-triangle: 7 x 128 bit; 28 mod 32 keeps 8 different records conflict-free
__global__ void __launch_bounds__(TPB, 4) k_eval(ConvertParams P) {
    __shared__ unsigned long long s_off[TPB + 1];
    __shared__ __align__(16) unsigned s_rec[TPB / 32][32 * EV2_REC];
    __shared__ unsigned s_pre[TPB / 32][33];
    TriCursor cur = locate(P, s_off);
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned* rec = s_rec[w] + lane * EV2_REC;
    f3 A{0.f, 0.f, 0.f}, B = A, C = A;
    int mn[3] = {0, 0, 0}, mx[3] = {-1, -1, -1};
    bool ok = false;
    if (cur.valid) {
        const float* p = tri_ptr(P, cur.tri);
        make_subtri(ld3(p), ld3(p + 3), ld3(p + 6), P.vs, cur.local, A, B, C);
        ok = subtri_box(A, B, C, P.inv_vs, P.band, mn, mx);
    }
    const int dx = ok ? mx[0] - mn[0] + 1 : 0, dy = ok ? mx[1] - mn[1] + 1 : 0, dz = ok ? mx[2] - mn[2] + 1 : 0;
    const bool wide = dx > 9 || dy > 9 || dz > 9;
    unsigned* vals = reinterpret_cast<unsigned*>(P.values);
    if (__any_sync(0xFFFFFFFFu, wide)) {  // rare (large narrow bands): per-thread loops over the whole box
        if (!ok) return;
        PtdTri T{A.x, A.y, A.z, B.x, B.y, B.z, C.x, C.y, C.z, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        ptd_tri_setup(T);
        for (int bx = mn[0] >> 3; bx <= (mx[0] >> 3); ++bx)
            for (int by = mn[1] >> 3; by <= (mx[1] >> 3); ++by)
                for (int bz = mn[2] >> 3; bz <= (mx[2] >> 3); ++bz) {
                    const unsigned slot = hash_lookup(P, bs_brick_key(bx, by, bz));
                    if (slot == 0xFFFFFFFFu) continue;  // brick not kept on this rank
                    unsigned* brick = vals + (size_t)slot * 512;
                    const int x0 = max(mn[0], bx << 3), x1 = min(mx[0], (bx << 3) + 7);
                    const int y0 = max(mn[1], by << 3), y1 = min(mx[1], (by << 3) + 7);
                    const int z0 = max(mn[2], bz << 3), z1 = min(mx[2], (bz << 3) + 7);
                    for (int x = x0; x <= x1; ++x)
                        for (int y = y0; y <= y1; ++y) {
                            PtdCol K;
                            ptd_col_setup(T, xmul((float)x, P.vs), xmul((float)y, P.vs), K);
                            unsigned* line = brick + ((x & 7) << 6) + ((y & 7) << 3);
                            for (int z = z0; z <= z1; ++z)
                                atomicMin(line + (z & 7), __float_as_uint(ptd_eval2(T, K, xmul((float)z, P.vs))));  // squared distance: >= 0 or NaN (NaN bits sort above the sentinel)
                        }
                }
        return;
    }
    // park: vertices, box origin, dims + the multiply-shift constant for "/ dy", brick slots
    {
        uint4* r4 = reinterpret_cast<uint4*>(rec);
        r4[0] = make_uint4(__float_as_uint(A.x), __float_as_uint(A.y), __float_as_uint(A.z), __float_as_uint(B.x));
        r4[1] = make_uint4(__float_as_uint(B.y), __float_as_uint(B.z), __float_as_uint(C.x), __float_as_uint(C.y));
        r4[2] = make_uint4(__float_as_uint(C.z), (unsigned)mn[0], (unsigned)mn[1], (unsigned)mn[2]);
        const unsigned magic = dy > 0 ? 65536u / (unsigned)dy + 1u : 0u;  // (q * magic) >> 16 == q / dy for q < 81, dy <= 9
        r4[3] = make_uint4((unsigned)dy | ((unsigned)dz << 8), magic, 0u, 0u);
        unsigned sl[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) sl[k] = 0xFFFFFFFFu;
        if (ok) {
            const int bx0 = mn[0] >> 3, by0 = mn[1] >> 3, bz0 = mn[2] >> 3;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int bx = bx0 + (k & 1), by = by0 + ((k >> 1) & 1), bz = bz0 + (k >> 2);
                const bool touched = bx <= (mx[0] >> 3) && by <= (mx[1] >> 3) && bz <= (mx[2] >> 3);
                if (touched) sl[k] = hash_lookup(P, bs_brick_key(bx, by, bz));
            }
        }
        r4[4] = make_uint4(sl[0], sl[1], sl[2], sl[3]);
        r4[5] = make_uint4(sl[4], sl[5], sl[6], sl[7]);
    }
    // exclusive prefix of column counts
    const unsigned ncol = (unsigned)(dx * dy);
    unsigned inc = ncol;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if ((int)lane >= o) inc += t; }
    s_pre[w][lane + 1] = inc;
    if (lane == 0) s_pre[w][0] = 0;
    __syncwarp();
    const unsigned total = s_pre[w][32];
    for (unsigned c = lane; c < total; c += 32) {
        unsigned lo = 0, hi = 32;  // owner: last s with pre[s] <= c
#pragma unroll
        for (int it = 0; it < 5; ++it) { const unsigned mid = (lo + hi) >> 1; if (s_pre[w][mid] <= c) lo = mid; else hi = mid; }
        const unsigned* r = s_rec[w] + lo * EV2_REC;
        const uint4* r4 = reinterpret_cast<const uint4*>(r);
        const uint4 q0 = r4[0], q1 = r4[1], q2 = r4[2], q3 = r4[3];
        PtdTri T{__uint_as_float(q0.x), __uint_as_float(q0.y), __uint_as_float(q0.z), __uint_as_float(q0.w), __uint_as_float(q1.x), __uint_as_float(q1.y),
                 __uint_as_float(q1.z), __uint_as_float(q1.w), __uint_as_float(q2.x), 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        ptd_tri_setup(T);
        const unsigned q = c - s_pre[w][lo];
        const unsigned ddy = q3.x & 255u, ddz = q3.x >> 8;
        const unsigned xi = (q * q3.y) >> 16;
        const unsigned yi = q - xi * ddy;
        const int x = (int)q2.y + (int)xi, y = (int)q2.z + (int)yi, z0 = (int)q2.w;
        const int kx = (x >> 3) - ((int)q2.y >> 3), ky = (y >> 3) - ((int)q2.z >> 3);
        const unsigned slot_lo = r[16 + kx + 2 * ky], slot_hi = r[20 + kx + 2 * ky];
        PtdCol K;
        ptd_col_setup(T, xmul((float)x, P.vs), xmul((float)y, P.vs), K);
        const unsigned line_off = ((x & 7) << 6) + ((y & 7) << 3);
        // a brick this rank does not keep (sharded runs) is replaced by a scratch brick: no test inside the loop
        unsigned* const line_lo = (slot_lo != 0xFFFFFFFFu ? vals + (size_t)slot_lo * 512 : P.dummy_brick) + line_off;
        unsigned* const line_hi = (slot_hi != 0xFFFFFFFFu ? vals + (size_t)slot_hi * 512 : P.dummy_brick) + line_off;
        const int zsplit = ((z0 >> 3) + 1) << 3;  // first z of the upper brick
        for (unsigned zi = 0; zi < ddz; ++zi) {
            const int z = z0 + (int)zi;
            const float d2 = ptd_eval2(T, K, xmul((float)z, P.vs));
            atomicMin((z >= zsplit ? line_hi : line_lo) + (z & 7), __float_as_uint(d2));  // d2 >= 0 or NaN (NaN bits sort above the sentinel)
        }
    }
}

// ---- blocked lattice edges (sign propagation on closed meshes, bs_signprop.cu) ----------------------------------------------
// For every lattice edge (p, p + e_a) that an ORIGINAL triangle meets -- or comes within a rounding margin of -- set the
// edge's bit in blk[brick(p)][a]: bit ((y&7)<<3 | (z&7)) of word a*8 + (x&7). Two active lattice neighbours whose edge
// is not blocked see the same winding number, exactly. The test is a conservative rasterisation in fp64: for each axis a
// the triangle is projected along a, every lattice column (u, v) inside the projection inflated by eps is intersected
// with the triangle's plane, and the lattice edges overlapping [z - dz, z + dz] are blocked (dz grows with the slope; a
// projection thinner than eps -- an edge-on triangle -- blocks its whole extent along a). Lattice positions are the f32
// products idx * vs every other stage uses, widened exactly to double; vertices are the f32 inputs: the predicate itself
// only rounds at 2^-53. eps = 16 * 2^-24 * (longest edge) + 2e-6 * vs is there for the REFERENCE's arithmetic: its f32
// solid angles (aabb_tree.rs:582-615) resolve the side of a triangle only down to a few 2^-24 of the triangle's size, so
// a voxel that close to the surface loses all six edges and is evaluated on its own, like the reference does.
constexpr unsigned RASTER_SMALL = 256;  // columns a single thread walks; larger projections go to k_block_edges_big
__device__ __forceinline__ double lat(int i, float vs) { return (double)__fmul_rn((float)i, vs); }
// smallest index whose lattice position is >= x / largest index whose lattice position is <= x (positions are monotone in the index)
__device__ __forceinline__ int lat_first(double x, double ivs, float vs) {
    int i = (int)ceil(x * ivs);
    while (lat(i - 1, vs) >= x) --i;
    while (lat(i, vs) < x) ++i;
    return i;
}
__device__ __forceinline__ int lat_last(double x, double ivs, float vs) {
    int i = (int)floor(x * ivs);
    while (lat(i + 1, vs) <= x) ++i;
    while (lat(i, vs) > x) --i;
    return i;
}
struct RasterTri {  // one triangle projected along axis A: coordinates (u, v, z) = (axis A+1, A+2, A)
    double U0, V0, Z0, U1, V1, Z1, U2, V2, Z2;
    double eps, umin, umax, vmin, vmax, zlo, zhi;       // zlo / zhi: extent along the axis, already widened by eps
    double du0, dv0, du1, dv1, du2, dv2, m0, m1, m2;   // edge vectors of the projection, m_k = eps * |edge k|
    double s, gu, gv, dz, ivs;                          // orientation, plane gradient, half-width of the z interval, 1 / vs
    double lu, lv, lm, lU, lV;                          // thin projections: direction, margin and start of the longest edge
    int iu0, iv0, nu, nv; bool thin, ok;
};
template <int A> __device__ __forceinline__ void raster_setup(const ConvertParams& P, size_t t, RasterTri& R) {
    const float* p = tri_ptr(P, t);
    constexpr int UA = (A + 1) % 3, VA = (A + 2) % 3;
    R.U0 = (double)p[UA]; R.V0 = (double)p[VA]; R.Z0 = (double)p[A];
    R.U1 = (double)p[3 + UA]; R.V1 = (double)p[3 + VA]; R.Z1 = (double)p[3 + A];
    R.U2 = (double)p[6 + UA]; R.V2 = (double)p[6 + VA]; R.Z2 = (double)p[6 + A];
    const double e1u = R.U1 - R.U0, e1v = R.V1 - R.V0, e1z = R.Z1 - R.Z0, e2u = R.U2 - R.U0, e2v = R.V2 - R.V0, e2z = R.Z2 - R.Z0;
    const double e3u = R.U2 - R.U1, e3v = R.V2 - R.V1, e3z = R.Z2 - R.Z1;
    const double l2 = fmax(e1u * e1u + e1v * e1v + e1z * e1z, fmax(e2u * e2u + e2v * e2v + e2z * e2z, e3u * e3u + e3v * e3v + e3z * e3z));
    R.ok = l2 < 1.0e60;  // finite
    R.nu = R.nv = 0;
    if (!R.ok) return;
    R.ivs = 1.0 / (double)P.vs;
    R.eps = 16.0 * 5.9604644775390625e-8 * ((double)sqrtf((float)l2) * 1.0001 + 1.0e-300) + 2.0e-6 * (double)P.vs;
    R.umin = fmin(R.U0, fmin(R.U1, R.U2)); R.umax = fmax(R.U0, fmax(R.U1, R.U2));
    R.vmin = fmin(R.V0, fmin(R.V1, R.V2)); R.vmax = fmax(R.V0, fmax(R.V1, R.V2));
    R.zlo = fmin(R.Z0, fmin(R.Z1, R.Z2)) - R.eps; R.zhi = fmax(R.Z0, fmax(R.Z1, R.Z2)) + R.eps;
    const double lim = 1048000.0;
    if (!(R.umin * R.ivs > -lim && R.umax * R.ivs < lim && R.vmin * R.ivs > -lim && R.vmax * R.ivs < lim && R.zlo * R.ivs > -lim && R.zhi * R.ivs < lim)) { R.ok = false; return; }  // outside the index range: k_mark reports it
    // exactly the lattice lines whose f32 position lies in [min - eps, max + eps]
    const int u0 = lat_first(R.umin - R.eps, R.ivs, P.vs), u1 = lat_last(R.umax + R.eps, R.ivs, P.vs);
    const int v0 = lat_first(R.vmin - R.eps, R.ivs, P.vs), v1 = lat_last(R.vmax + R.eps, R.ivs, P.vs);
    R.iu0 = u0; R.iv0 = v0; R.nu = max(0, u1 - u0 + 1); R.nv = max(0, v1 - v0 + 1);
    R.du0 = e1u; R.dv0 = e1v; R.du1 = e3u; R.dv1 = e3v; R.du2 = -e2u; R.dv2 = -e2v;
    const double len0 = (double)sqrtf((float)(e1u * e1u + e1v * e1v)) * 1.0001, len1 = (double)sqrtf((float)(e3u * e3u + e3v * e3v)) * 1.0001, len2 = (double)sqrtf((float)(e2u * e2u + e2v * e2v)) * 1.0001;
    R.m0 = R.eps * len0; R.m1 = R.eps * len1; R.m2 = R.eps * len2;
    const double nu_ = e1v * e2z - e1z * e2v, nv_ = e1z * e2u - e1u * e2z, nz_ = e1u * e2v - e1v * e2u;  // (u, v, z) components of e1 x e2
    R.thin = !(fabs(nz_) > R.eps * (len0 + len1 + len2));  // projected height below ~eps (or NaN): edge-on
    R.s = nz_ >= 0.0 ? 1.0 : -1.0;
    if (!R.thin) {
        const double inz = 1.0 / nz_;
        R.gu = -nu_ * inz; R.gv = -nv_ * inz; R.dz = R.eps * (2.0 + fabs(R.gu) + fabs(R.gv));
    } else {  // the projection is (within eps) the segment of its longest edge
        R.gu = R.gv = R.dz = 0.0;
        if (len0 >= len1 && len0 >= len2) { R.lu = R.du0; R.lv = R.dv0; R.lm = 3.0 * R.m0; R.lU = R.U0; R.lV = R.V0; }
        else if (len1 >= len2) { R.lu = R.du1; R.lv = R.dv1; R.lm = 3.0 * R.m1; R.lU = R.U1; R.lV = R.V1; }
        else { R.lu = R.du2; R.lv = R.dv2; R.lm = 3.0 * R.m2; R.lU = R.U2; R.lV = R.V2; }
    }
}
template <int A> __device__ __forceinline__ void raster_block(const ConvertParams& P, unsigned long long* blk, int iu, int iv, int k, unsigned long long& last_key, unsigned& last_slot) {
    int x, y, z;
    if (A == 0) { x = k; y = iu; z = iv; } else if (A == 1) { y = k; z = iu; x = iv; } else { z = k; x = iu; y = iv; }
    const int bx = x >> 3, by = y >> 3, bz = z >> 3;
    if (bx < BS_BRICK_MIN || bx > BS_BRICK_MAX || by < BS_BRICK_MIN || by > BS_BRICK_MAX || bz < BS_BRICK_MIN || bz > BS_BRICK_MAX) return;
    const unsigned long long key = bs_brick_key(bx, by, bz);
    if (key != last_key) { last_key = key; last_slot = hash_lookup(P, key); }
    if (last_slot == 0xFFFFFFFFu) return;  // no such brick (or not kept on this rank): the edge has no active lower end
    atomicOr(blk + (size_t)last_slot * 24 + A * 8 + (x & 7), 1ull << (((y & 7) << 3) | (z & 7)));
}
template <int A> __device__ __forceinline__ void raster_column(const ConvertParams& P, unsigned long long* blk, const RasterTri& R, int iu, int iv, double pu, unsigned long long& last_key, unsigned& last_slot) {
    const double pv = lat(iv, P.vs);
    double lo = R.zlo, hi = R.zhi;
    if (!R.thin) {
        if (R.s * (R.du0 * (pv - R.V0) - R.dv0 * (pu - R.U0)) < -R.m0) return;  // outside the projection inflated by eps
        if (R.s * (R.du1 * (pv - R.V1) - R.dv1 * (pu - R.U1)) < -R.m1) return;
        if (R.s * (R.du2 * (pv - R.V2) - R.dv2 * (pu - R.U2)) < -R.m2) return;
        const double zc = R.Z0 + R.gu * (pu - R.U0) + R.gv * (pv - R.V0);
        const double a = fmax(lo, zc - R.dz), b = fmin(hi, zc + R.dz);
        if (a <= b) { lo = a; hi = b; }  // (else: numerically impossible; stay with the full extent)
    } else if (fabs(R.lu * (pv - R.lV) - R.lv * (pu - R.lU)) > R.lm) return;
    // lattice edges [k, k + 1] that overlap [lo, hi]: from the last index at or below lo (minus one when lo is a lattice
    // position itself) to the last index at or below hi
    const int k0 = lat_first(lo, R.ivs, P.vs) - 1, k1 = lat_last(hi, R.ivs, P.vs);
    for (int k = k0; k <= k1; ++k) raster_block<A>(P, blk, iu, iv, k, last_key, last_slot);
}
template <int A> __device__ __forceinline__ void raster_small(const ConvertParams& P, unsigned long long* blk, size_t t, unsigned long long g, unsigned long long* big_list, unsigned* n_big, unsigned big_cap) {
    RasterTri R;
    raster_setup<A>(P, t, R);
    if (!R.ok) return;  // non-finite input never passes the closedness test; out-of-range indices were reported by k_mark
    if ((unsigned long long)R.nu * (unsigned long long)R.nv > RASTER_SMALL) {
        const unsigned i = atomicAdd(n_big, 1u);
        if (i < big_cap) { big_list[i] = g; return; }  // (a full list: this thread walks the columns itself)
    }
    unsigned long long last_key = BS_KEY_INVALID; unsigned last_slot = 0xFFFFFFFFu;
    for (int i = 0; i < R.nu; ++i) {
        const int iu = R.iu0 + i;
        const double pu = lat(iu, P.vs);
        for (int j = 0; j < R.nv; ++j) raster_column<A>(P, blk, R, iu, R.iv0 + j, pu, last_key, last_slot);
    }
}
template <int A> __device__ __forceinline__ void raster_big(const ConvertParams& P, unsigned long long* blk, size_t t) {
    RasterTri R;
    raster_setup<A>(P, t, R);
    if (!R.ok) return;
    const unsigned long long ncols = (unsigned long long)R.nu * (unsigned long long)R.nv;
    unsigned long long last_key = BS_KEY_INVALID; unsigned last_slot = 0xFFFFFFFFu;
    for (unsigned long long c = threadIdx.x; c < ncols; c += blockDim.x) {
        const int iu = R.iu0 + (int)(c / (unsigned)R.nv), iv = R.iv0 + (int)(c % (unsigned)R.nv);
        raster_column<A>(P, blk, R, iu, iv, lat(iu, P.vs), last_key, last_slot);
    }
}
// (Measured and rejected, r2: parking the set-up projections in shared memory and dealing their lattice columns out to the
// lanes of the warp, as k_eval does -- 2.33 ms against 2.05 ms for this per-thread walk.)
// 8 CTAs / SM (64 registers, 216 B of spills): the kernel waits on its loads -- the triangle, the brick hash -- and ran at 12 warps
// per SM with the 106 registers the fp64 set-up would like (2.05 ms; 5 / 6 / 8 CTAs: 1.74 / 1.55 / 1.35 ms).
__global__ void __launch_bounds__(128, 8) k_block_edges(ConvertParams P, unsigned long long* blk, unsigned long long* big_list, unsigned* n_big, unsigned big_cap) {
    // one thread per (axis, triangle); the axis is the SLOW index, so a warp runs one instantiation of the rasteriser
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g >= P.n_tris * 3) return;
    const size_t t = g % P.n_tris; const int a = (int)(g / P.n_tris);
    if (P.use_clip) {  // sharded: the triangle's box (plus slack) misses every brick this rank keeps
        const float* p = tri_ptr(P, t);
        for (int d = 0; d < 3; ++d) {
            const float lo = fminf(p[d], fminf(p[3 + d], p[6 + d])), hi = fmaxf(p[d], fmaxf(p[3 + d], p[6 + d]));
            if (ceilf(hi * P.inv_vs) + 3.0f < (float)P.clip_mn[d] || floorf(lo * P.inv_vs) - 3.0f > (float)P.clip_mx[d]) return;
        }
    }
    if (a == 0) raster_small<0>(P, blk, t, g, big_list, n_big, big_cap);
    else if (a == 1) raster_small<1>(P, blk, t, g, big_list, n_big, big_cap);
    else raster_small<2>(P, blk, t, g, big_list, n_big, big_cap);
}
// projections of more than RASTER_SMALL columns: one CTA per (triangle, axis), threads stride over the columns
__global__ void __launch_bounds__(256) k_block_edges_big(ConvertParams P, unsigned long long* blk, const unsigned long long* big_list, const unsigned* n_big, unsigned big_cap) {
    const unsigned n = min(*n_big, big_cap);
    for (unsigned i = blockIdx.x; i < n; i += gridDim.x) {
        const unsigned long long g = big_list[i];
        const size_t t = (size_t)(g % P.n_tris); const int a = (int)(g / P.n_tris);
        if (a == 0) raster_big<0>(P, blk, t); else if (a == 1) raster_big<1>(P, blk, t); else raster_big<2>(P, blk, t);
    }
}

__global__ void k_fill_slots(const unsigned long long* sorted_keys, size_t n, unsigned long long* table_keys, unsigned* table_slots, unsigned mask, unsigned* derr) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long key = sorted_keys[i];
    unsigned h = (unsigned)hash64(key) & mask;
    int probe = 0;
    while (table_keys[h] != key) { h = (h + 1) & mask; if (++probe > MAX_PROBE) { atomicOr(derr, BS_DERR_PROBE); return; } }
    table_slots[h] = (unsigned)i;
}

__device__ __forceinline__ long long find_sorted(const unsigned long long* keys, size_t n, unsigned long long k) {
    size_t lo = 0, hi = n;
    while (lo < hi) { size_t mid = (lo + hi) >> 1; if (keys[mid] < k) lo = mid + 1; else hi = mid; }
    return (lo < n && keys[lo] == k) ? (long long)lo : -1;
}
// keep[] = slab bricks [lo, hi) and their 26 neighbours
__global__ void k_mark_slab(const unsigned long long* __restrict__ keys, size_t n, size_t lo, size_t hi, unsigned char* keep) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= (hi - lo) * 27) return;
    const size_t b = lo + i / 27; const int d = (int)(i % 27);
    if (d == 13) { keep[b] = 1; return; }
    int bx, by, bz; bs_key_brick(keys[b], bx, by, bz);
    bx += d / 9 - 1; by += (d / 3) % 3 - 1; bz += d % 3 - 1;
    if (bx < BS_BRICK_MIN || bx > BS_BRICK_MAX || by < BS_BRICK_MIN || by > BS_BRICK_MAX || bz < BS_BRICK_MIN || bz > BS_BRICK_MAX) return;
    const long long j = find_sorted(keys, n, bs_brick_key(bx, by, bz));
    if (j >= 0) keep[j] = 1;
}
// weight of sorted brick i = touches(i) + mean touches; W[i] = inclusive prefix. bounds[r] = first i with W[i] >= r * W_total / world
__global__ void k_brick_touches(const unsigned long long* __restrict__ keys, size_t n, const unsigned long long* __restrict__ table_keys, const unsigned* __restrict__ table_counts, unsigned mask, unsigned long long* touches, unsigned* derr) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long key = keys[i];
    unsigned h = (unsigned)hash64(key) & mask;
    int probe = 0;
    while (table_keys[h] != key) { h = (h + 1) & mask; if (++probe > MAX_PROBE) { atomicOr(derr, BS_DERR_PROBE); touches[i] = 0; return; } }
    touches[i] = table_counts[h];
}
// cost model of a brick for the sign stage (which dominates): t = sub-triangle boxes touching it, m = mean t.
// w = m + t + t^2 / m: linear for ordinary bricks, quadratic for bricks under dense slivers, where every voxel is "near"
// every triangle (measured on the UV-sphere poles of config 5: t = 33 m costs ~650 ordinary bricks; t + m gave 17)
__global__ void k_brick_weights(const unsigned long long* __restrict__ touches, const unsigned long long* __restrict__ C_touch, size_t n, unsigned long long* w) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long m = C_touch[n - 1] / n + 1, t = touches[i];
    w[i] = m + t + t * t / m;
}
__global__ void k_slab_bounds(const unsigned long long* __restrict__ C /*inclusive scan of the weights*/, size_t n, int world, unsigned long long* bounds /*world + 1*/) {
    const int r = threadIdx.x;
    if (r > world) return;
    if (r == 0) { bounds[0] = 0; return; }
    if (r == world) { bounds[world] = n; return; }
    const unsigned long long target = C[n - 1] / (unsigned long long)world * (unsigned long long)r;
    size_t lo = 0, hi = n;  // first i with C[i] >= target
    while (lo < hi) { const size_t mid = (lo + hi) >> 1; if (C[mid] < target) lo = mid + 1; else hi = mid; }
    bounds[r] = lo;
}
__global__ void k_key_bounds(const unsigned long long* __restrict__ keys, size_t n, int* bounds /*min xyz, max xyz in voxels*/) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
    if (i < n) { int b[3]; bs_key_brick(keys[i], b[0], b[1], b[2]); for (int d = 0; d < 3; ++d) { lo[d] = b[d] << 3; hi[d] = (b[d] << 3) + 7; } }
    for (int d = 0; d < 3; ++d) {
        for (int o = 16; o; o >>= 1) { lo[d] = min(lo[d], __shfl_xor_sync(0xFFFFFFFFu, lo[d], o)); hi[d] = max(hi[d], __shfl_xor_sync(0xFFFFFFFFu, hi[d], o)); }
        if ((threadIdx.x & 31) == 0) { atomicMin(bounds + d, lo[d]); atomicMax(bounds + 3 + d, hi[d]); }
    }
}
__global__ void k_owned(const unsigned long long* __restrict__ all_keys, size_t n_all, size_t lo, size_t hi, const unsigned long long* __restrict__ kept, size_t n_kept, unsigned char* owned) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n_kept) return;
    const long long j = find_sorted(all_keys, n_all, kept[i]);
    owned[i] = (j >= (long long)lo && j < (long long)hi) ? 1 : 0;
}

// ---- sharded runs: coarse cut of the key space and the triangles a rank needs -----------------------------------------------
// Coarse cell = brick key >> 8: one x-plane of bricks inside a 128^3 node (8 x 128 x 128 voxels); a contiguous range of
// coarse keys is a contiguous range of brick keys, so slabs cut at coarse keys keep "concatenate the ranks' outputs in rank
// order == the single-GPU output". Every rank derives the same cut from a histogram of the triangles' sub-triangle counts
// per coarse cell (integer atomics: deterministic), then works only on the triangles whose box -- inflated by one brick of
// halo, the sub-triangle box rules and the band -- can reach its range.
__device__ __forceinline__ unsigned long long coarse_key_of_voxel(int x, int y, int z) { return bs_brick_key(x >> 3, y >> 3, z >> 3) >> 8; }
__global__ void k_coarse_hist(const float* __restrict__ tris, size_t n_tris, float vs, float inv_vs, unsigned long long* tkeys, unsigned long long* tw, unsigned mask, int* flags) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= n_tris) return;
    const float* p = tris + 9 * t;
    const f3 p1 = ld3(p), p2 = ld3(p + 3), p3 = ld3(p + 6);
    const float n = num_subs_of(p1, p2, p3, vs);
    unsigned long long w = 1;
    if (n >= 2.0f) { const double nd = fmin((double)n, 4.0e9); w = (unsigned long long)(nd * nd); } else if (n != n) return;
    const float cx = (p1.x + p2.x + p3.x) * (1.0f / 3.0f) * inv_vs, cy = (p1.y + p2.y + p3.y) * (1.0f / 3.0f) * inv_vs, cz = (p1.z + p2.z + p3.z) * (1.0f / 3.0f) * inv_vs;
    if (!(fabsf(cx) < 1.0e6f && fabsf(cy) < 1.0e6f && fabsf(cz) < 1.0e6f)) return;  // k_mark reports range errors
    const unsigned long long key = coarse_key_of_voxel((int)floorf(cx), (int)floorf(cy), (int)floorf(cz));
    // neighbouring triangles mostly share a cell: one table update per distinct key of the (active part of the) warp
    const unsigned act = __activemask(), peers = __match_any_sync(act, key), leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
    unsigned long long sum = 0;
    for (unsigned m = peers; m; m &= m - 1) sum += __shfl_sync(peers, w, __ffs(m) - 1);
    if (lane != leader) return;
    unsigned h = (unsigned)hash64(key) & mask;
    for (int probe = 0; probe < 4096; ++probe) {
        const unsigned long long cur = tkeys[h];
        bool mine = cur == key;
        if (!mine && cur == BS_KEY_INVALID) { const unsigned long long prev = atomicCAS(&tkeys[h], BS_KEY_INVALID, key); mine = prev == BS_KEY_INVALID || prev == key; }
        if (mine) { atomicAdd(&tw[h], sum); return; }
        h = (h + 1) & mask;
    }
    flags[0] = 1;  // table full: the caller falls back to an unbalanced but valid cut
}
// bounds[r] = coarse key at which rank r starts (bounds[0] = 0, bounds[world] = KEY_INVALID): first key whose inclusive
// weight prefix reaches r / world of the total
__global__ void k_coarse_cut(const unsigned long long* __restrict__ keys /*sorted*/, const unsigned long long* __restrict__ C /*inclusive scan of the weights*/, size_t n, int world, unsigned long long* bounds) {
    const int r = threadIdx.x;
    if (r > world) return;
    if (r == 0) { bounds[0] = 0; return; }
    if (r == world || n == 0) { bounds[r] = BS_KEY_INVALID; return; }
    const unsigned long long target = C[n - 1] / (unsigned long long)world * (unsigned long long)r;
    size_t lo = 0, hi = n;  // first i with C[i] >= target
    while (lo < hi) { const size_t mid = (lo + hi) >> 1; if (C[mid] < target) lo = mid + 1; else hi = mid; }
    bounds[r] = lo < n ? keys[lo] : BS_KEY_INVALID;
}
__global__ void k_select_tris(const float* __restrict__ tris, size_t n_tris, float inv_vs, int margin /*voxels*/, unsigned long long klo, unsigned long long khi, unsigned char* keep) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= n_tris) return;
    const float* p = tris + 9 * t;
    int lo[3], hi[3]; bool ok = true;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float a = floorf(fminf(p[d], fminf(p[3 + d], p[6 + d])) * inv_vs), b = ceilf(fmaxf(p[d], fmaxf(p[3 + d], p[6 + d])) * inv_vs);
        if (!(a > -1.0e6f && b < 1.0e6f)) { ok = false; lo[d] = hi[d] = 0; }  // NaN / out of range: keep it, k_mark reports
        else { lo[d] = ((int)a - margin) >> 3; hi[d] = ((int)b + margin) >> 3; }  // brick coordinates
        lo[d] = max(lo[d], BS_BRICK_MIN); hi[d] = min(hi[d], BS_BRICK_MAX);
    }
    bool hit = !ok;
    if (ok) {
        // coarse cells = (brick x, node y, node z): x per brick, y / z per 16 bricks
        const long long cells = (long long)(hi[0] - lo[0] + 1) * ((hi[1] >> 4) - (lo[1] >> 4) + 1) * ((hi[2] >> 4) - (lo[2] >> 4) + 1);
        if (cells > 256) hit = true;  // a huge triangle: every rank takes it
        else
            for (int bx = lo[0]; bx <= hi[0] && !hit; ++bx)
                for (int ny = lo[1] >> 4; ny <= (hi[1] >> 4) && !hit; ++ny)
                    for (int nz = lo[2] >> 4; nz <= (hi[2] >> 4) && !hit; ++nz) {
                        const unsigned long long k = bs_brick_key(bx, ny * 16, nz * 16) >> 8;
                        hit = k >= klo && k < khi;
                    }
    }
    keep[t] = hit ? 1 : 0;
}
__global__ void k_count_below(const unsigned long long* __restrict__ keys, size_t n, unsigned long long b0, unsigned long long b1, unsigned long long* out /*[2]*/) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    unsigned c0 = 0, c1 = 0;
    if (i < n) { const unsigned long long k = keys[i]; if (k != BS_KEY_INVALID) { c0 = k < b0; c1 = k < b1; } }
    c0 = __popc(__ballot_sync(0xFFFFFFFFu, c0)); c1 = __popc(__ballot_sync(0xFFFFFFFFu, c1));
    if ((threadIdx.x & 31) == 0) { if (c0) atomicAdd(out, (unsigned long long)c0); if (c1) atomicAdd(out + 1, (unsigned long long)c1); }
}

struct NotEmptyKey { __device__ bool operator()(unsigned long long k) const { return k != BS_KEY_INVALID; } };

__global__ void k_counts(const float* values, const unsigned long long* masks, size_t n_bricks, unsigned long long* out /*[2]*/) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;  // one thread per mask word
    unsigned long long a = 0, neg = 0;
    if (i < n_bricks * 8) {
        unsigned long long m = masks[i];
        a = __popcll(m);
        const float* v = values + i * 64;
        while (m) { int b = __ffsll((long long)m) - 1; m &= m - 1; if (__float_as_uint(v[b]) >> 31) ++neg; }
    }
    typedef cub::BlockReduce<unsigned long long, TPB> BR;
    __shared__ typename BR::TempStorage t1;
    unsigned long long sa = BR(t1).Sum(a);
    __syncthreads();
    unsigned long long sn = BR(t1).Sum(neg);
    if (threadIdx.x == 0) { if (sa) atomicAdd(out, sa); if (sn) atomicAdd(out + 1, sn); }
}

}  // namespace

extern "C" bs_status bs_volume_counts(const bs_volume* v, size_t* n_bricks, size_t* n_active, size_t* n_negative, size_t* n_tiles) {
    if (!v || !v->ctx) return BS_ERR_INVALID;
    bs_context* ctx = v->ctx;
    BS_ENTER(ctx);
    unsigned long long* d = nullptr; unsigned long long h[2] = {0, 0};
    BS_TRY(bs_alloc(ctx, &d, 2));
    BS_CUDA(ctx, cudaMemsetAsync(d, 0, 16, ctx->stream));
    if (v->n_bricks) bs_count_launch(), k_counts<<<bs_blocks(v->n_bricks * 8, TPB), TPB, 0, ctx->stream>>>(v->values, v->masks, v->n_bricks, d);
    BS_TRY(bs_fetch(ctx, h, d, 16));
    BS_TRY(bs_sync(ctx));
    bs_free(ctx, d);
    if (n_bricks) *n_bricks = v->n_bricks;
    if (n_active) *n_active = h[0];
    if (n_negative) *n_negative = h[1];
    if (n_tiles) *n_tiles = v->n_tiles8 + v->n_tiles128;
    return BS_OK;
}

bs_status bs_convert_impl(bs_context* ctx, const float* d_tris, size_t n_tris, float voxel_size, int64_t band, int rank, int world, bs_volume** out, bs_convert_plan* plan) {
    const bool planned = plan && plan->valid && plan->world == world && (int)plan->bounds.size() == world + 1;  // a later slab of the same mesh
    cudaStream_t st = ctx->stream;
    bs_marks_begin(ctx);
    BS_CUDA(ctx, cudaMemsetAsync(ctx->d_err, 0, sizeof(unsigned), st));
    const size_t n_mesh = n_tris;  // triangles of the whole mesh (the sign stage needs all of them)
    int* d_flags = nullptr;
    BS_TRY(bs_alloc(ctx, &d_flags, 2));
    BS_CUDA(ctx, cudaMemsetAsync(d_flags, 0, 2 * sizeof(int), st));
    // closed mesh? (bs_signprop.cu) -- enqueued here, read at the next synchronisation
    bs_closed_check chk; chk.pending = false; chk.closed = false; chk.exact = false; chk.d_sums = nullptr; chk.d_bad = nullptr;
    ctx->mesh_closed = false;
    if (ctx->sign_propagation && !planned) BS_TRY(bs_mesh_closed_begin(ctx, d_tris, n_tris, &chk));
    // 0. sharded: cut the coarse key space at equal sub-triangle weight, keep the triangles that can reach this rank's range
    unsigned* d_tri_ids = nullptr; unsigned long long klo = 0, khi = BS_KEY_INVALID;
    if (world > 1 && planned) { klo = plan->bounds[rank]; khi = plan->bounds[rank + 1]; }
    else if (world > 1) {
        const unsigned tcap = 1u << 17;  // coarse cells that can be occupied: the table is sorted whole, keep it small
        unsigned long long *d_tk = nullptr, *d_tw = nullptr, *d_sk = nullptr, *d_sw = nullptr, *d_C = nullptr, *d_bounds = nullptr; size_t* d_nsel = nullptr; void* d_tmp0 = nullptr; size_t tmp0 = 0, tmp1 = 0, tmp2 = 0;
        BS_TRY(bs_alloc(ctx, &d_tk, (size_t)tcap)); BS_TRY(bs_alloc(ctx, &d_tw, (size_t)tcap)); BS_TRY(bs_alloc(ctx, &d_sk, (size_t)tcap)); BS_TRY(bs_alloc(ctx, &d_sw, (size_t)tcap));
        BS_TRY(bs_alloc(ctx, &d_C, (size_t)tcap)); BS_TRY(bs_alloc(ctx, &d_bounds, (size_t)world + 1)); BS_TRY(bs_alloc(ctx, &d_nsel, 1));
        BS_CUDA(ctx, cudaMemsetAsync(d_tk, 0xFF, tcap * sizeof(unsigned long long), st));
        BS_CUDA(ctx, cudaMemsetAsync(d_tw, 0, tcap * sizeof(unsigned long long), st));
        bs_count_launch(), k_coarse_hist<<<bs_blocks(n_tris, TPB), TPB, 0, st>>>(d_tris, n_tris, voxel_size, 1.0f / voxel_size, d_tk, d_tw, tcap - 1, d_flags);
        // sort the (key, weight) pairs: empty slots (key = all ones, weight 0) go to the end
        cub::DeviceRadixSort::SortPairs(nullptr, tmp0, d_tk, d_sk, d_tw, d_sw, (int)tcap, 0, 64, st);
        cub::DeviceScan::InclusiveSum(nullptr, tmp1, d_sw, d_C, (int)tcap, st);
        cub::DeviceSelect::If(nullptr, tmp2, d_tk, d_sk, d_nsel, (int)tcap, NotEmptyKey(), st);
        BS_TRY(bs_alloc(ctx, (char**)&d_tmp0, std::max(tmp0, std::max(tmp1, tmp2))));
        cub::DeviceSelect::If(d_tmp0, tmp2, d_tk, d_sk, d_nsel, (int)tcap, NotEmptyKey(), st);  // (only for the count of occupied cells)
        size_t n_cells = 0; int hflag[2] = {0, 0};
        BS_TRY(bs_fetch(ctx, &n_cells, d_nsel, sizeof(size_t)));
        cub::DeviceRadixSort::SortPairs(d_tmp0, tmp0, d_tk, d_sk, d_tw, d_sw, (int)tcap, 0, 64, st);
        cub::DeviceScan::InclusiveSum(d_tmp0, tmp1, d_sw, d_C, (int)tcap, st);
        BS_TRY(bs_fetch(ctx, hflag, d_flags, sizeof(hflag)));
        BS_TRY(bs_sync(ctx));
        bs_count_launch(), k_coarse_cut<<<1, 64, 0, st>>>(d_sk, d_C, hflag[0] ? 0 : n_cells, world, d_bounds);
        std::vector<unsigned long long> hb((size_t)world + 1);
        BS_TRY(bs_fetch(ctx, hb.data(), d_bounds, hb.size() * sizeof(unsigned long long)));
        BS_TRY(bs_sync(ctx));
        if (hflag[0]) { BS_CUDA(ctx, cudaMemsetAsync(d_flags, 0, 2 * sizeof(int), st)); }  // (table full: rank 0 takes everything; still exact)
        klo = hb[rank]; khi = hb[rank + 1];
        if (plan) { plan->bounds = hb; plan->world = world; }
        bs_free(ctx, d_tmp0); bs_free(ctx, d_tk); bs_free(ctx, d_tw); bs_free(ctx, d_sk); bs_free(ctx, d_sw); bs_free(ctx, d_C); bs_free(ctx, d_bounds); bs_free(ctx, d_nsel);
    }
    if (world > 1) {
        size_t* d_nsel = nullptr; void* d_tmp0 = nullptr; size_t tmp0 = 0;
        BS_TRY(bs_alloc(ctx, &d_nsel, 1));
        // triangles whose inflated box reaches [klo, khi): one brick of halo + sub-triangle box slack (2) + band + 1
        unsigned char* d_keep = nullptr;
        BS_TRY(bs_alloc(ctx, &d_keep, n_tris)); BS_TRY(bs_alloc(ctx, &d_tri_ids, n_tris));
        bs_count_launch(), k_select_tris<<<bs_blocks(n_tris, TPB), TPB, 0, st>>>(d_tris, n_tris, 1.0f / voxel_size, 11 + (int)band, klo, khi, d_keep);
        tmp0 = 0;
        cub::DeviceSelect::Flagged(nullptr, tmp0, cub::CountingInputIterator<unsigned>(0), d_keep, d_tri_ids, d_nsel, (int)n_tris, st);
        BS_TRY(bs_alloc(ctx, (char**)&d_tmp0, tmp0));
        cub::DeviceSelect::Flagged(d_tmp0, tmp0, cub::CountingInputIterator<unsigned>(0), d_keep, d_tri_ids, d_nsel, (int)n_tris, st);
        size_t n_list = 0;
        BS_TRY(bs_fetch(ctx, &n_list, d_nsel, sizeof(size_t)));
        BS_TRY(bs_sync(ctx));
        bs_free(ctx, d_tmp0); bs_free(ctx, d_keep); bs_free(ctx, d_nsel);
        n_tris = n_list;  // from here on: this rank's triangle list
        bs_mark(ctx, "shard_select_ms");
    }
    // 1. per-triangle sub-triangle counts + surface area in voxel^2 (sizing only)
    unsigned long long *d_counts = nullptr, *d_offsets = nullptr; double* d_area = nullptr;
    BS_TRY(bs_alloc(ctx, &d_counts, n_tris + 1)); BS_TRY(bs_alloc(ctx, &d_offsets, n_tris + 1));
    BS_TRY(bs_alloc(ctx, &d_area, 1));
    unsigned long long* d_neval = nullptr;
    BS_TRY(bs_alloc(ctx, &d_neval, 1));
    BS_CUDA(ctx, cudaMemsetAsync(d_area, 0, sizeof(double), st));
    BS_CUDA(ctx, cudaMemsetAsync(d_counts + n_tris, 0, sizeof(unsigned long long), st));
    if (n_tris) bs_count_launch(), k_tri_counts<<<bs_blocks(n_tris, TPB), TPB, 0, st>>>(d_tris, d_tri_ids, n_tris, voxel_size, d_counts, d_area);
    void* d_tmp = nullptr; size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_counts, d_offsets, n_tris + 1, st);
    BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp_bytes));
    cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_counts, d_offsets, n_tris + 1, st);
    unsigned long long total = 0; double area_vox = 0.0;
    BS_TRY(bs_fetch(ctx, &total, d_offsets + n_tris, sizeof(total)));
    BS_TRY(bs_fetch(ctx, &area_vox, d_area, sizeof(double)));
    BS_TRY(bs_sync(ctx));
    ctx->mesh_closed = planned ? (ctx->sign_propagation && plan->closed) : (ctx->sign_propagation && bs_mesh_closed_finish(ctx, &chk));
    if (plan && !planned) { plan->closed = ctx->mesh_closed; plan->world = world; plan->valid = world == 1 || (int)plan->bounds.size() == world + 1; }
    bs_free(ctx, d_tmp); bs_free(ctx, d_counts); bs_free(ctx, d_area);
    bs_mark(ctx, "subdivide_count_ms");
    if (total == 0 && world > 1) {  // nothing of the mesh reaches this rank's range: an empty slab
        bs_free(ctx, d_offsets); bs_free(ctx, d_flags); bs_free(ctx, d_neval); bs_free(ctx, d_tri_ids);
        bs_volume* ev = bs_volume_new(ctx, voxel_size);
        bs_status es = bs_volume_alloc_bricks(ev, 0);
        if (es == BS_OK) es = bs_alloc(ctx, &ev->owned, (size_t)1);
        if (es != BS_OK) { bs_volume_free(ev); return es; }
        ev->n_owned = 0;
        bs_marks_end(ctx);
        bs_stat_add(ctx, "n_tris", (double)n_mesh); bs_stat_add(ctx, "n_bricks", 0.0); bs_stat_add(ctx, "n_bricks_owned", 0.0); bs_stat_add(ctx, "n_active", 0.0);
        *out = ev;
        return BS_OK;
    }
    if (total == 0) { bs_free(ctx, d_offsets); bs_free(ctx, d_flags); bs_free(ctx, d_neval); bs_marks_end(ctx); return BS_ERR_EMPTY_MESH; }  // convert -> None (:58-60)
    if (total > (1ull << 40)) { bs_free(ctx, d_offsets); bs_free(ctx, d_flags); bs_free(ctx, d_neval); return bs_fail(ctx, BS_ERR_RANGE, "%llu sub-triangles: voxel size too small for this mesh", total); }

    ConvertParams P;
    P.tris = d_tris; P.n_tris = n_tris; P.tri_ids = d_tri_ids; P.offsets = d_offsets; P.total = total;
    P.vs = voxel_size; P.inv_vs = 1.0f / voxel_size; P.band = (int)band; P.flags = d_flags; P.values = nullptr; P.n_eval = d_neval; P.table_counts = nullptr; P.use_clip = 0; P.derr = ctx->d_err;
    P.cta_start = nullptr; P.dummy_brick = nullptr;
    const unsigned grid = (unsigned)((total + TPB - 1) / TPB);
    unsigned* d_cta_start = nullptr; unsigned* d_dummy = nullptr;
    BS_TRY(bs_alloc(ctx, &d_cta_start, (size_t)grid)); BS_TRY(bs_alloc(ctx, &d_dummy, (size_t)512));
    bs_count_launch(), k_cta_starts<<<bs_blocks(grid, TPB), TPB, 0, st>>>(P, grid, d_cta_start);
    P.cta_start = d_cta_start; P.dummy_brick = d_dummy;

    // 2. mark touched bricks in a hash set; sized from the surface area, doubled on overflow
    const double bw = (double)(2 * band + 1);
    double est = 0.15 * area_vox * bw + 8192.0;
    if (est > 6.0e8) est = 6.0e8;
    size_t cap = 1; while ((double)cap < 2.0 * est) cap <<= 1;
    unsigned long long* d_table_keys = nullptr; unsigned* d_table_slots = nullptr; unsigned* d_table_counts = nullptr;
    unsigned long long* d_keys = nullptr; size_t n_all = 0; unsigned long long n_eval = 0;
    for (;;) {
        BS_TRY(bs_alloc(ctx, &d_table_keys, cap));
        BS_CUDA(ctx, cudaMemsetAsync(d_table_keys, 0xFF, cap * sizeof(unsigned long long), st));
        { BS_TRY(bs_alloc(ctx, &d_table_counts, cap)); BS_CUDA(ctx, cudaMemsetAsync(d_table_counts, 0, cap * sizeof(unsigned), st)); }
        P.table_counts = d_table_counts;
        BS_CUDA(ctx, cudaMemsetAsync(d_neval, 0, sizeof(unsigned long long), st));
        P.table_keys = d_table_keys; P.table_slots = nullptr; P.table_mask = (unsigned)(cap - 1);
        bs_count_launch(), k_mark<<<grid, TPB, 0, st>>>(P);
        int flags[2];
        BS_TRY(bs_fetch(ctx, flags, d_flags, sizeof(flags)));
        BS_TRY(bs_fetch(ctx, &n_eval, d_neval, sizeof(n_eval)));
        BS_TRY(bs_sync(ctx));
        if (flags[1]) { bs_free(ctx, d_table_keys); bs_free(ctx, d_offsets); bs_free(ctx, d_flags); return bs_fail(ctx, BS_ERR_RANGE, "voxel index outside [-2^20, 2^20)"); }
        if (!flags[0]) break;
        bs_free(ctx, d_table_keys); bs_free(ctx, d_table_counts); d_table_counts = nullptr;
        BS_CUDA(ctx, cudaMemsetAsync(d_flags, 0, 2 * sizeof(int), st));
        cap <<= 1;
        if (cap > (1ull << 31)) { bs_free(ctx, d_offsets); bs_free(ctx, d_flags); return bs_fail(ctx, BS_ERR_RANGE, "brick hash set overflow"); }
    }
    bs_mark(ctx, "mark_bricks_ms");
    // 3. compact + sort the keys (ascending key == the reference's leaf visit order)
    {
        unsigned long long* d_sel = nullptr; size_t* d_nsel = nullptr;
        BS_TRY(bs_alloc(ctx, &d_sel, cap)); BS_TRY(bs_alloc(ctx, &d_nsel, 1));
        tmp_bytes = 0;
        cub::DeviceSelect::If(nullptr, tmp_bytes, d_table_keys, d_sel, d_nsel, cap, NotEmptyKey(), st);
        BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp_bytes));
        cub::DeviceSelect::If(d_tmp, tmp_bytes, d_table_keys, d_sel, d_nsel, cap, NotEmptyKey(), st);
        BS_TRY(bs_fetch(ctx, &n_all, d_nsel, sizeof(size_t)));
        BS_TRY(bs_sync(ctx));
        bs_free(ctx, d_tmp); bs_free(ctx, d_nsel);
        BS_TRY(bs_alloc(ctx, &d_keys, n_all));
        tmp_bytes = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, d_sel, d_keys, n_all, 0, 54, st);
        BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp_bytes));
        cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, d_sel, d_keys, n_all, 0, 54, st);
        bs_free(ctx, d_tmp); bs_free(ctx, d_sel);
    }
    bs_volume* vol = bs_volume_new(ctx, voxel_size);
    bs_status s = BS_OK;
    const size_t n_total = n_all;
    if (world > 1) {
        // brick-slab sharding: this rank owns the bricks whose coarse key lies in [klo, khi) -- a contiguous slab [lo, hi) of
        // the sorted brick list -- and keeps, as read-only halo, the 26 neighbours of its bricks (extraction needs +1 for MC,
        // -1..+1 for DC). The local list also holds bricks further out (the triangle selection is conservative): dropped here.
        size_t lo, hi;
        {
            unsigned long long* d_lh = nullptr; unsigned long long h_lh[2] = {0, 0};
            BS_TRY(bs_alloc(ctx, &d_lh, 2));
            BS_CUDA(ctx, cudaMemsetAsync(d_lh, 0, 2 * sizeof(unsigned long long), st));
            const unsigned long long b0 = klo << 8, b1 = khi == BS_KEY_INVALID ? BS_KEY_INVALID : (khi << 8);
            if (n_all) bs_count_launch(), k_count_below<<<bs_blocks(n_all, TPB), TPB, 0, st>>>(d_keys, n_all, b0, b1, d_lh);
            BS_TRY(bs_fetch(ctx, h_lh, d_lh, sizeof(h_lh)));
            BS_TRY(bs_sync(ctx));
            bs_free(ctx, d_lh);
            lo = (size_t)h_lh[0]; hi = (size_t)h_lh[1];
            if (hi < lo) hi = lo;
        }
        unsigned char* d_keep = nullptr; unsigned long long* d_kept = nullptr; size_t* d_nk = nullptr; size_t n_kept = 0;
        BS_TRY(bs_alloc(ctx, &d_keep, n_all)); BS_TRY(bs_alloc(ctx, &d_kept, n_all)); BS_TRY(bs_alloc(ctx, &d_nk, 1));
        BS_CUDA(ctx, cudaMemsetAsync(d_keep, 0, n_all, st));
        if (hi > lo) bs_count_launch(), k_mark_slab<<<bs_blocks((hi - lo) * 27, TPB), TPB, 0, st>>>(d_keys, n_all, lo, hi, d_keep);
        tmp_bytes = 0;
        cub::DeviceSelect::Flagged(nullptr, tmp_bytes, d_keys, d_keep, d_kept, d_nk, n_all, st);
        BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp_bytes));
        cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, d_keys, d_keep, d_kept, d_nk, n_all, st);
        BS_TRY(bs_fetch(ctx, &n_kept, d_nk, sizeof(size_t)));
        BS_TRY(bs_sync(ctx));
        bs_free(ctx, d_tmp); bs_free(ctx, d_nk); bs_free(ctx, d_keep);
        s = bs_volume_alloc_bricks(vol, n_kept);
        if (s == BS_OK) s = bs_alloc(ctx, &vol->owned, n_kept);
        if (s != BS_OK) { bs_volume_free(vol); return s; }
        BS_CUDA(ctx, cudaMemcpyAsync(vol->keys, d_kept, n_kept * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
        if (n_kept) bs_count_launch(), k_owned<<<bs_blocks(n_kept, TPB), TPB, 0, st>>>(d_keys, n_all, lo, hi, vol->keys, n_kept, vol->owned);
        bs_free(ctx, d_kept); bs_free(ctx, d_keys);
        n_all = n_kept;
        vol->n_owned = hi - lo;
    } else {
        s = bs_volume_alloc_bricks(vol, n_all);
        if (s != BS_OK) { bs_volume_free(vol); return s; }
        BS_CUDA(ctx, cudaMemcpyAsync(vol->keys, d_keys, n_all * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
        bs_free(ctx, d_keys);
    }
    BS_TRY(bs_alloc(ctx, &d_table_slots, cap));
    BS_CUDA(ctx, cudaMemsetAsync(d_table_slots, 0xFF, cap * sizeof(unsigned), st));
    bs_count_launch(), k_fill_slots<<<bs_blocks(n_all, TPB), TPB, 0, st>>>((const unsigned long long*)vol->keys, n_all, d_table_keys, d_table_slots, (unsigned)(cap - 1), ctx->d_err);
    BS_CUDA(ctx, cudaMemsetAsync(vol->values, 0x7F, n_all * 512 * sizeof(float), st));
    bs_mark(ctx, "sort_bricks_ms");
    // 4. distances
    P.table_slots = d_table_slots; P.values = vol->values;
    P.use_clip = 0;
    // (sharded runs work on this rank's triangle list: no further clipping)
    bs_count_launch(), k_eval<<<grid, TPB, 0, st>>>(P);
    bs_mark(ctx, "udf_ms");
    // closed mesh: lattice edges met by a triangle (sign propagation, bs_signprop.cu); needs the brick hash, so it runs here
    unsigned long long* d_blk = nullptr;
    if (ctx->mesh_closed && n_all) {
        const unsigned big_cap = 1u << 22;
        unsigned long long* d_big = nullptr; unsigned* d_nbig = nullptr;
        BS_TRY(bs_alloc(ctx, &d_blk, n_all * 24)); BS_TRY(bs_alloc(ctx, &d_big, (size_t)big_cap)); BS_TRY(bs_alloc(ctx, &d_nbig, 1));
        BS_CUDA(ctx, cudaMemsetAsync(d_blk, 0, n_all * 24 * sizeof(unsigned long long), st));
        BS_CUDA(ctx, cudaMemsetAsync(d_nbig, 0, sizeof(unsigned), st));
        bs_count_launch(), k_block_edges<<<bs_blocks(n_tris * 3, 128), 128, 0, st>>>(P, d_blk, d_big, d_nbig, big_cap);
        bs_count_launch(), k_block_edges_big<<<(unsigned)ctx->sm_count * 8, TPB, 0, st>>>(P, d_blk, d_big, d_nbig, big_cap);
        bs_free(ctx, d_big); bs_free(ctx, d_nbig);
        bs_mark(ctx, "sign_block_edges_ms");
    }
    // per-brick "touches" (sub-triangle boxes that hit the brick): the sign stage runs the densest bricks first
    unsigned long long* d_touch_kept = nullptr;
    BS_TRY(bs_alloc(ctx, &d_touch_kept, n_all));
    if (n_all) bs_count_launch(), k_brick_touches<<<bs_blocks(n_all, TPB), TPB, 0, st>>>(vol->keys, n_all, d_table_keys, d_table_counts, (unsigned)(cap - 1), d_touch_kept, ctx->d_err);
    bs_free(ctx, d_table_keys); bs_free(ctx, d_table_slots); bs_free(ctx, d_table_counts); bs_free(ctx, d_offsets); bs_free(ctx, d_flags); bs_free(ctx, d_neval);
    bs_free(ctx, d_cta_start); bs_free(ctx, d_dummy);
    // 5. signs + masks
    s = bs_sign_impl(ctx, d_tris, n_mesh, vol, d_touch_kept, d_blk);  // winding numbers see the whole mesh
    bs_free(ctx, d_touch_kept); bs_free(ctx, d_blk); bs_free(ctx, d_tri_ids);
    if (s != BS_OK) { bs_volume_free(vol); return s; }
    BS_CUDA(ctx, cudaGetLastError());
    unsigned derr = 0;
    BS_TRY(bs_fetch(ctx, &derr, ctx->d_err, sizeof(unsigned)));
    bs_marks_end(ctx);  // synchronises
    if (derr) {  // a kernel ran out of a fixed-size resource: an error, never a silently wrong volume
        bs_volume_free(vol);
        return bs_fail(ctx, BS_ERR_RANGE, "device limit exceeded:%s%s", (derr & BS_DERR_STACK) ? " winding-number traversal stack" : "", (derr & BS_DERR_PROBE) ? " brick hash probe length" : "");
    }
    bs_stat_add(ctx, "n_tris", (double)n_mesh);
    bs_stat_add(ctx, "n_tris_local", (double)n_tris);
    bs_stat_add(ctx, "n_sub", (double)total);
    bs_stat_add(ctx, "n_bricks", (double)n_all);
    bs_stat_add(ctx, "n_bricks_total", (double)n_total);
    bs_stat_add(ctx, "n_bricks_owned", (double)vol->n_owned);
    bs_stat_add(ctx, "n_eval", (double)n_eval);
    if (ctx->count_work) {
        bs_stat_add(ctx, "fwn_visits", ctx->fwn_counts[0]); bs_stat_add(ctx, "fwn_far", ctx->fwn_counts[1]);
        bs_stat_add(ctx, "fwn_exact_tris", ctx->fwn_counts[2]); bs_stat_add(ctx, "fwn_voxels", ctx->fwn_counts[3]);
        bs_stat_add(ctx, "fwn_warp_visits", ctx->fwn_counts[4]); bs_stat_add(ctx, "fwn_traversals", ctx->fwn_counts[5]);
        bs_stat_add(ctx, "fwn_brick_visits", ctx->fwn_counts[6]); bs_stat_add(ctx, "fwn_brick_hoisted", ctx->fwn_counts[7]); bs_stat_add(ctx, "fwn_brick_roots", ctx->fwn_counts[8]);
    }
    bs_stat_add(ctx, "area_vox", area_vox);
    *out = vol;
    return BS_OK;
}
