// Marching Cubes 33 iso-surface extraction on sorted bricks (sm_100a).
// Replaces MarchingCubesMesher::mesh (src/voxel/meshing/marching_cubes.rs:43-63): ComputeEdgeIntersections
// (:1013-1121), Cube::from_voxel (:1161-1181), handle_cube (:72-285), add_faces (:287-317), the face /
// interior tests (:336-916), compute_c_vertex (:918-938); tables from lookup_table.rs via mc33_tables.h.
//
// The reference builds three auxiliary sparse grids of edge intersections and then walks every cell with
// 8 root-to-leaf lookups. Here one CTA owns one brick: it stages the brick plus its +x/+y/+z halo (9^3
// values, active bits) in shared memory, one thread classifies one cell, and edge intersections are
// recomputed from the staged values (same f32 expression, so the same bits). Output order is the
// reference's: bricks in key order (== leaf visit order), cells x-major inside a brick, triangles in table
// order -- reproduced with a count pass, an exclusive scan over bricks and an emit pass that block-scans
// the per-cell counts. All arithmetic that feeds a vertex or a branch uses the non-contracting x* helpers.
#include "bs_common.cuh"
#include "mc33_tables.h"
#include <cub/cub.cuh>
#include <thrust/iterator/transform_iterator.h>
#include <cfloat>
#include <algorithm>

namespace {

constexpr int MC_TPB = 256;
constexpr float MIN_ABS = 1e-6f;  // MIN_ABS_VERTEX_VALUE (marching_cubes.rs:1123)

__constant__ unsigned char c_iav[96] = MC33_IAV_PERM_INIT;
__constant__ signed char c_edge_v1[12] = {0, 1, 3, 0, 4, 5, 7, 4, 0, 1, 2, 3};
__constant__ signed char c_edge_v2[12] = {1, 2, 2, 3, 5, 6, 6, 7, 4, 5, 6, 7};
__constant__ signed char c_edge_dir[12] = {0, 1, 0, 1, 0, 1, 0, 1, 2, 2, 2, 2};
__constant__ signed char c_corner[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};

struct VolView {
    const unsigned long long* keys; const float* values; const unsigned long long* masks; size_t n;
    const unsigned char* owned;  // nullptr = all
    const int* nbr;              // 8 per brick: index of the brick at (+dx,+dy,+dz), bit0 = x, bit1 = y, bit2 = z; -1 = absent
    const unsigned long long* t8k; const float* t8v; size_t nt8;
    const unsigned long long* t128k; const float* t128v; size_t nt128;
};

__device__ __forceinline__ long long find_key(const unsigned long long* keys, size_t n, unsigned long long k) {
    size_t lo = 0, hi = n;
    while (lo < hi) { size_t mid = (lo + hi) >> 1; if (keys[mid] < k) lo = mid + 1; else hi = mid; }
    return (lo < n && keys[lo] == k) ? (long long)lo : -1;
}

struct Cell {
    float c[8];       // clamped corner values (Cube::from_voxel)
    int ox, oy, oz;   // global index of corner 0
    const signed char* T;
    int cs, cf;
    f3 v12;
};

__device__ __forceinline__ bool test_face(const Cell& q, int face) {  // :378-401
    float a, b, c, d;
    switch (face < 0 ? -face : face) {
        case 1: a = q.c[0]; b = q.c[4]; c = q.c[5]; d = q.c[1]; break;
        case 2: a = q.c[1]; b = q.c[5]; c = q.c[6]; d = q.c[2]; break;
        case 3: a = q.c[2]; b = q.c[6]; c = q.c[7]; d = q.c[3]; break;
        case 4: a = q.c[3]; b = q.c[7]; c = q.c[4]; d = q.c[0]; break;
        case 5: a = q.c[0]; b = q.c[3]; c = q.c[2]; d = q.c[1]; break;
        case 6: a = q.c[4]; b = q.c[7]; c = q.c[6]; d = q.c[5]; break;
        default: return false;
    }
    float val = xsub(xmul(a, c), xmul(b, d));
    if (fabsf(val) < FLT_EPSILON) return face >= 0;
    return xmul(xmul((float)face, a), val) >= 0.f;
}
__device__ __forceinline__ int interior_ambiguity(const Cell& q, int amb_face, int face_i) {  // :474-538
    const float f = (float)face_i;
    const bool p17 = xmul(q.c[1], f) > 0.f && xmul(q.c[7], f) > 0.f, p06 = xmul(q.c[0], f) > 0.f && xmul(q.c[6], f) > 0.f;
    const bool p35 = xmul(q.c[3], f) > 0.f && xmul(q.c[5], f) > 0.f, p24 = xmul(q.c[2], f) > 0.f && xmul(q.c[4], f) > 0.f;
    int e = 0;
    switch (amb_face) {
        case 1: case 3: if (p17) e = 4; if (p06) e = 5; if (p35) e = 6; if (p24) e = 7; break;
        case 2: case 4: if (p17) e = 0; if (p24) e = 1; if (p35) e = 2; if (p06) e = 3; break;
        case 5: case 6: case 0: if (p06) e = 8; if (p17) e = 9; if (p24) e = 10; if (p35) e = 11; break;
        default: break;
    }
    return e;
}
__device__ __forceinline__ int iav(const Cell& q, int edge) {  // interior_ambiguity_verification :540-916
    const unsigned char* p = c_iav + 8 * edge;
    const float A0 = q.c[p[0]], A1 = q.c[p[1]], B0 = q.c[p[2]], B1 = q.c[p[3]], C0 = q.c[p[4]], C1 = q.c[p[5]], D0 = q.c[p[6]], D1 = q.c[p[7]];
    const float dA = xsub(A1, A0), dB = xsub(B1, B0), dC = xsub(C1, C0), dD = xsub(D1, D0);
    const float a = xsub(xmul(dA, dC), xmul(dB, dD));
    const float b = xsub(xsub(xadd(xmul(C0, dA), xmul(A0, dC)), xmul(D0, dB)), xmul(B0, dD));
    if (a > 0.f) return 1;
    const float t = xdiv(-b, xmul(2.0f, a));
    if (t < 0.f || t > 1.f) return 1;
    const float at = xadd(A0, xmul(dA, t)), bt = xadd(B0, xmul(dB, t)), ct = xadd(C0, xmul(dC, t)), dt = xadd(D0, xmul(dD, t));
    const float verify = xsub(xmul(at, ct), xmul(bt, dt));
    if (verify > 0.f) return 0;
    if (verify < 0.f) return 1;
    return 0;
}
#define TB(name, cfg, i) ((int)q.T[MC33_OFF_##name + (cfg) * MC33_ROW_##name + (i)])
__device__ bool test_interior(const Cell& q, int face) {  // :403-472
    switch (q.cs) {
        case 4: return (iav(q, interior_ambiguity(q, 1, face)) + iav(q, interior_ambiguity(q, 2, face)) + iav(q, interior_ambiguity(q, 5, face))) != 0;
        case 6: return iav(q, interior_ambiguity(q, abs(TB(TEST_6, q.cf, 0)), face)) != 0;
        case 7: { const int s = -face; return (iav(q, interior_ambiguity(q, 1, s)) + iav(q, interior_ambiguity(q, 2, s)) + iav(q, interior_ambiguity(q, 5, s))) != 0; }
        case 10: return iav(q, interior_ambiguity(q, abs(TB(TEST_10, q.cf, 0)), face)) != 0;
        case 12: return (iav(q, interior_ambiguity(q, abs(TB(TEST_12, q.cf, 0)), face)) + iav(q, interior_ambiguity(q, abs(TB(TEST_12, q.cf, 1)), face))) != 0;
        default: return false;
    }
}
__device__ bool interior_test_case13(const Cell& q) {  // :336-376
    const float* c = q.c;
    const float d01 = xsub(c[0], c[1]), d76 = xsub(c[7], c[6]), d45 = xsub(c[4], c[5]), d32 = xsub(c[3], c[2]);
    const float a = xsub(xmul(d01, d76), xmul(d45, d32));
    const float b = xsub(xsub(xadd(xmul(c[6], d01), xmul(c[1], d76)), xmul(c[2], d45)), xmul(c[5], d32));
    const float cc = xsub(xmul(c[1], c[6]), xmul(c[5], c[2]));
    const float delta = xsub(xmul(b, b), xmul(xmul(4.0f, a), cc));
    const float sq = xsqrt(delta), a2 = xadd(a, a);
    const float t1 = xdiv(xadd(-b, sq), a2), t2 = xdiv(xsub(-b, sq), a2);
    if (t1 < 1.f && t1 > 0.f && t2 < 1.f && t2 > 0.f) {
        float xy[4];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float t = k ? t2 : t1;
            const float at = xadd(c[1], xmul(d01, t)), bt = xadd(c[5], xmul(d45, t)), ct = xadd(c[6], xmul(d76, t)), dt = xadd(c[2], xmul(d32, t));
            const float den = xsub(xsub(xadd(at, ct), bt), dt);
            xy[2 * k] = xdiv(xsub(at, dt), den); xy[2 * k + 1] = xdiv(xsub(at, bt), den);
        }
        return !(xy[0] < 1.f && xy[0] > 0.f && xy[2] < 1.f && xy[2] > 0.f && xy[1] < 1.f && xy[1] > 0.f && xy[3] < 1.f && xy[3] > 0.f);
    }
    return true;
}

// handle_cube (:72-285): pick the tiling row for this cell. Returns the blob offset, its length in `len`
// and whether the interior c-vertex is needed.
#define ROW1(name, cfg) (len = MC33_ROW_##name, MC33_OFF_##name + (cfg) * MC33_ROW_##name)
#define ROW2(name, cfg, sub) (len = MC33_ROW_##name, MC33_OFF_##name + ((cfg) * MC33_SUB_##name + (sub)) * MC33_ROW_##name)
__device__ int select_tiling(const Cell& q, int& len, bool& need_c, bool& stale) {
    const int cf = q.cf;
    need_c = false; stale = false; len = 0;
    switch (q.cs) {
        case 1: return ROW1(TILING_1, cf);
        case 2: return ROW1(TILING_2, cf);
        case 3: return test_face(q, q.T[MC33_OFF_TEST_3 + cf]) ? ROW1(TILING_3_2, cf) : ROW1(TILING_3_1, cf);
        case 4: return test_interior(q, q.T[MC33_OFF_TEST_4 + cf]) ? ROW1(TILING_4_1, cf) : ROW1(TILING_4_2, cf);
        case 5: return ROW1(TILING_5, cf);
        case 6:
            if (test_face(q, TB(TEST_6, cf, 0))) return ROW1(TILING_6_2, cf);
            if (test_interior(q, TB(TEST_6, cf, 1))) return ROW1(TILING_6_1_1, cf);
            stale = true;  // tiling 6.1.2 uses the c-vertex but the reference does not compute it here (marching_cubes.rs:103-111)
            return ROW1(TILING_6_1_2, cf);
        case 7: {
            int sub = 0;
            if (test_face(q, TB(TEST_7, cf, 0))) sub += 1;
            if (test_face(q, TB(TEST_7, cf, 1))) sub += 2;
            if (test_face(q, TB(TEST_7, cf, 2))) sub += 4;
            switch (sub) {
                case 0: return ROW1(TILING_7_1, cf);
                case 1: return ROW2(TILING_7_2, cf, 0);
                case 2: return ROW2(TILING_7_2, cf, 1);
                case 3: need_c = true; return ROW2(TILING_7_3, cf, 0);
                case 4: return ROW2(TILING_7_2, cf, 2);
                case 5: need_c = true; return ROW2(TILING_7_3, cf, 1);
                case 6: need_c = true; return ROW2(TILING_7_3, cf, 2);
                default: return test_interior(q, TB(TEST_7, cf, 3)) ? ROW1(TILING_7_4_1, cf) : ROW1(TILING_7_4_2, cf);
            }
        }
        case 8: return ROW1(TILING_8, cf);
        case 9: return ROW1(TILING_9, cf);
        case 10:
            if (test_face(q, TB(TEST_10, cf, 0))) {
                if (test_face(q, TB(TEST_10, cf, 1))) return test_interior(q, -TB(TEST_10, cf, 2)) ? ROW1(TILING_10_1_1_, cf) : ROW1(TILING_10_1_2, 5 - cf);
                need_c = true; return ROW1(TILING_10_2, cf);
            }
            if (test_face(q, TB(TEST_10, cf, 1))) { need_c = true; return ROW1(TILING_10_2_, cf); }
            return test_interior(q, TB(TEST_10, cf, 2)) ? ROW1(TILING_10_1_1, cf) : ROW1(TILING_10_1_2, cf);
        case 11: return ROW1(TILING_11, cf);
        case 12:
            if (test_face(q, TB(TEST_12, cf, 0))) {
                if (test_face(q, TB(TEST_12, cf, 1))) return test_interior(q, -TB(TEST_12, cf, 2)) ? ROW1(TILING_12_1_1_, cf) : ROW1(TILING_12_1_2, 23 - cf);
                need_c = true; return ROW1(TILING_12_2, cf);
            }
            if (test_face(q, TB(TEST_12, cf, 1))) { need_c = true; return ROW1(TILING_12_2_, cf); }
            return test_interior(q, TB(TEST_12, cf, 2)) ? ROW1(TILING_12_1_1, cf) : ROW1(TILING_12_1_2, cf);
        case 13: {
            int sub = 0;
            for (int i = 0; i < 6; ++i) if (test_face(q, TB(TEST_13, cf, i))) sub += 1 << i;
            const int sc = q.T[MC33_OFF_SUB_CONFIG_13 + sub];
            if (sc == 0) return ROW1(TILING_13_1, cf);
            if (sc >= 1 && sc <= 6) return ROW2(TILING_13_2, cf, sc - 1);
            if (sc >= 7 && sc <= 18) { need_c = true; return ROW2(TILING_13_3, cf, sc - 7); }
            if (sc >= 19 && sc <= 22) { need_c = true; return ROW2(TILING_13_4, cf, sc - 19); }
            if (sc >= 23 && sc <= 26) {
                const bool in = interior_test_case13(q);
                if (cf == 0) return in ? ROW2(TILING_13_5_1, 0, sc - 23) : ROW2(TILING_13_5_2, 1, sc - 23);
                return in ? ROW2(TILING_13_5_1, 1, sc - 23) : ROW2(TILING_13_5_2, 0, sc - 23);
            }
            if (sc >= 27 && sc <= 38) { need_c = true; return ROW2(TILING_13_3_, cf, sc - 27); }
            if (sc >= 39 && sc <= 44) return ROW2(TILING_13_2_, cf, sc - 39);
            if (sc == 45) return ROW1(TILING_13_1_, cf);
            return 0;
        }
        case 14: return ROW1(TILING_14, cf);
        default: return 0;
    }
}

// MarchingCubesMesher::intersection (:320-334) with the x/y/z_int value recomputed (:1054-1079).
// Index-space point; false when the reference's aux grid has no entry (no sign change on the edge).
__device__ __forceinline__ bool edge_point(const Cell& q, int e, f3& out) {
    if (e == 12) { out = q.v12; return true; }
    const int v1 = c_edge_v1[e], v2 = c_edge_v2[e];
    const float a = q.c[v1], b = q.c[v2];
    if ((__float_as_uint(a) ^ __float_as_uint(b)) >> 31 == 0) return false;
    const float fa = fabsf(a), fb = fabsf(b);  // already >= MIN_ABS
    const float t = xdiv(fa, xadd(fa, fb));
    const int ix = q.ox + c_corner[v1][0], iy = q.oy + c_corner[v1][1], iz = q.oz + c_corner[v1][2];
    out = f3{(float)ix, (float)iy, (float)iz};
    const int dir = c_edge_dir[e];
    if (dir == 0) out.x = xadd(out.x, t); else if (dir == 1) out.y = xadd(out.y, t); else out.z = xadd(out.z, t);
    return true;
}
__device__ void compute_c_vertex(Cell& q) {  // :918-938
    f3 sum{0.f, 0.f, 0.f}; int count = 0;
    for (int e = 0; e < 12; ++e) { f3 p; if (edge_point(q, e, p)) { sum = xadd(sum, p); ++count; } }
    const float n = (float)count;
    q.v12 = f3{xdiv(sum.x, n), xdiv(sum.y, n), xdiv(sum.z, n)};
}

// add_faces (:287-317). WRITE = false: count only.
template <bool WRITE>
__device__ int emit_rows(const Cell& q, float vs, float* out, int row, int len) {
    int n = 0;
    for (int i = 0; i + 2 < len; i += 3) {
        const int e1 = q.T[row + i], e3 = q.T[row + i + 1], e2 = q.T[row + i + 2];
        f3 v1, v2, v3;
        if (!edge_point(q, e1, v1) || !edge_point(q, e2, v2) || !edge_point(q, e3, v3)) continue;
        v1 = xscale(v1, vs); v2 = xscale(v2, vs); v3 = xscale(v3, vs);
        if (xnorm2(xcross(xsub(v2, v1), xsub(v3, v1))) == 0.f) continue;  // Triangle3::is_degenerate
        if (WRITE) {
            float* o = out + 9 * n;
            o[0] = v1.x; o[1] = v1.y; o[2] = v1.z; o[3] = v2.x; o[4] = v2.y; o[5] = v2.z; o[6] = v3.x; o[7] = v3.y; o[8] = v3.z;
        }
        ++n;
    }
    return n;
}
template <bool WRITE>
__device__ int emit_cell(Cell& q, float vs, float* out) {  // cells outside the brick pass (tiles): no c-vertex carry
    int len; bool need_c, stale;
    const int row = select_tiling(q, len, need_c, stale);
    if (len == 0) return 0;
    if (need_c) compute_c_vertex(q);
    return emit_rows<WRITE>(q, vs, out, row, len);
}

// The reference keeps ONE c-vertex (`self.v12`) across cells and tiling 6.1.2 reads it without recomputing it, so a
// 6.1.2 cell emits the c-vertex of the latest earlier cell (in visit order) that computed one, or (0,0,0).
struct CarryV12 { float x, y, z; int valid; };
struct CarryOp { __device__ CarryV12 operator()(const CarryV12& a, const CarryV12& b) const { return b.valid ? b : a; } };

// Stage brick b and its +x/+y/+z halo: 9^3 values and active flags.
__device__ void stage_brick(const VolView& V, size_t b, float* s_val /*729*/, unsigned char* s_act /*729*/, long long* s_nb /*8*/, int* s_org /*3*/) {
    const unsigned tid = threadIdx.x;
    if (tid < 8) {
        if (tid == 0) { int bx, by, bz; bs_key_brick(V.keys[b], bx, by, bz); s_org[0] = bx << 3; s_org[1] = by << 3; s_org[2] = bz << 3; }
        s_nb[tid] = V.nbr[b * 8 + tid];  // precomputed: a binary search here would sit on every CTA's critical path
    }
    __syncthreads();
    for (unsigned i = tid; i < 729; i += blockDim.x) {
        const unsigned x = i / 81, y = (i / 9) % 9, z = i % 9;
        const unsigned nb = (x >> 3) | ((y >> 3) << 1) | ((z >> 3) << 2);
        const long long src = s_nb[nb];
        const unsigned off = ((x & 7) << 6) | ((y & 7) << 3) | (z & 7);
        float v = 0.f; unsigned char a = 0;
        if (src >= 0) {
            a = (V.masks[src * 8 + (off >> 6)] >> (off & 63)) & 1;
            v = V.values[src * 512 + off];
        } else if (V.nt8 | V.nt128) {
            // active tile fallback (only CSG creates tiles): TreeNode::at returns the tile value
            int bx = (s_org[0] >> 3) + (int)(x >> 3), by = (s_org[1] >> 3) + (int)(y >> 3), bz = (s_org[2] >> 3) + (int)(z >> 3);
            const unsigned long long k = bs_brick_key(bx, by, bz);
            long long t = V.nt8 ? find_key(V.t8k, V.nt8, k) : -1;
            if (t >= 0) { a = 1; v = V.t8v[t]; }
            else { t = V.nt128 ? find_key(V.t128k, V.nt128, k >> 12) : -1; if (t >= 0) { a = 1; v = V.t128v[t]; } }
        }
        s_val[i] = v; s_act[i] = a;
    }
    __syncthreads();
}

// Cube::from_voxel (:1161-1181) for the cell whose corner 0 is local (x,y,z)
__device__ __forceinline__ bool load_cell(Cell& q, const float* s_val, const unsigned char* s_act, const int* s_org, unsigned x, unsigned y, unsigned z, int& id) {
    id = 0;
    bool all = true;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const unsigned j = (x + c_corner[i][0]) * 81 + (y + c_corner[i][1]) * 9 + (z + c_corner[i][2]);
        all = all && s_act[j];
        float v = s_val[j];
        if (fabsf(v) < MIN_ABS) v = copysignf(MIN_ABS, v);
        if (v < 0.f) id |= 1 << i;
        q.c[i] = v;
    }
    q.ox = s_org[0] + (int)x; q.oy = s_org[1] + (int)y; q.oz = s_org[2] + (int)z;
    return all;
}

// `pos` maps a brick to its rank in the merged (bricks + active tiles) visit order; nullptr = identity.
// c-vertex carry: `writer[b]` receives the brick's last computed c-vertex (count pass); `incoming[b]` is the carry
// entering the brick (nullptr = not known yet: readers that need it are counted tentatively and flag the brick in
// `unresolved`); with `only_flagged` the count pass redoes just those bricks.
//
// Inside the CTA: (1) every thread looks at two cells and keeps those with all 8 corners active and a sign change;
// (2) the surviving cells (typically 60-120 of 512) are compacted in cell order, so the expensive MC33 logic runs on
// dense lanes; (3) per-cell triangle counts are scanned; (4) the emit pass writes triangles straight to their final
// place (no per-thread staging array).
constexpr int MC_CELLS_PER_THREAD = 512 / MC_TPB;
template <bool WRITE>
__global__ void __launch_bounds__(MC_TPB) k_mc(VolView V, const signed char* __restrict__ tables, float vs, const unsigned* __restrict__ pos,
                                                unsigned* item_counts, const unsigned long long* __restrict__ item_offsets, float* out,
                                                CarryV12* writer, const CarryV12* __restrict__ incoming, unsigned char* unresolved, int only_flagged, int* any_unresolved) {
    __shared__ float s_val[729];
    __shared__ unsigned char s_act[732];
    __shared__ long long s_nb[8];
    __shared__ int s_org[3];
    __shared__ unsigned short s_cells[512];      // compacted cell list: (cube id << 9) is kept separately
    __shared__ unsigned char s_ids[512];
    __shared__ unsigned s_wcount[MC_TPB / 32 * MC_CELLS_PER_THREAD + 1];
    __shared__ int s_maxw;
    __shared__ unsigned s_wmask[16];
    __shared__ float s_v12[512 * 3];
    const size_t b = blockIdx.x;
    const size_t item = pos ? pos[b] : b;
    if (V.owned && !V.owned[b]) { if (!WRITE && !only_flagged && threadIdx.x == 0) { item_counts[item] = 0; writer[b] = CarryV12{0.f, 0.f, 0.f, 0}; unresolved[b] = 0; } return; }  // halo brick of a sharded volume
    if (WRITE) { if (item_offsets[item + 1] == item_offsets[item]) return; }  // uniform per block
    if (!WRITE && only_flagged && !unresolved[b]) return;
    const unsigned tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_maxw = -1;
    if (tid < 16) s_wmask[tid] = 0;
    stage_brick(V, b, s_val, s_act, s_nb, s_org);
    // (1) + (2): candidate cells in cell order (cell c = leaf offset x<<6 | y<<3 | z of corner 0)
    unsigned bal[MC_CELLS_PER_THREAD]; int cid[MC_CELLS_PER_THREAD];
#pragma unroll
    for (int r = 0; r < MC_CELLS_PER_THREAD; ++r) {
        const unsigned c = r * MC_TPB + tid;
        const unsigned x = c >> 6, y = (c >> 3) & 7, z = c & 7;
        int id = 0; bool all = true;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const unsigned j = (x + c_corner[i][0]) * 81 + (y + c_corner[i][1]) * 9 + (z + c_corner[i][2]);
            all = all && s_act[j];
            id |= (int)(__float_as_uint(s_val[j]) >> 31) << i;  // clamping to +-1e-6 keeps the sign bit: id bit = sign bit
        }
        cid[r] = id;
        bal[r] = __ballot_sync(0xFFFFFFFFu, all && id != 0 && id != 255);
        if (lane == 0) s_wcount[r * (MC_TPB / 32) + w] = __popc(bal[r]);
    }
    __syncthreads();
    if (tid == 0) { unsigned acc = 0; for (int i = 0; i < MC_TPB / 32 * MC_CELLS_PER_THREAD; ++i) { const unsigned c = s_wcount[i]; s_wcount[i] = acc; acc += c; } s_wcount[MC_TPB / 32 * MC_CELLS_PER_THREAD] = acc; }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < MC_CELLS_PER_THREAD; ++r)
        if ((bal[r] >> lane) & 1) { const unsigned p = s_wcount[r * (MC_TPB / 32) + w] + __popc(bal[r] & ((1u << lane) - 1)); s_cells[p] = (unsigned short)(r * MC_TPB + tid); s_ids[p] = (unsigned char)cid[r]; }
    __syncthreads();
    const int n_act = (int)s_wcount[MC_TPB / 32 * MC_CELLS_PER_THREAD];
    // (3): MC33 on the compacted cells; n_act <= 512 -> up to MC_CELLS_PER_THREAD rounds, almost always one
    typedef cub::BlockScan<int, MC_TPB> Scan;
    __shared__ typename Scan::TempStorage tmp;
    unsigned long long running = 0;
    bool brick_unres = false;
    for (int base = 0; base < n_act || base == 0; base += MC_TPB) {
        const int j = base + (int)tid;
        Cell q; q.T = tables; q.v12 = f3{0.f, 0.f, 0.f};
        int row = 0, len = 0; bool need_c = false, stale = false;
        if (j < n_act) {
            const unsigned c = s_cells[j];
            int id;
            load_cell(q, s_val, s_act, s_org, c >> 6, (c >> 3) & 7, c & 7, id);
            q.cs = tables[MC33_OFF_CASES + 2 * id]; q.cf = tables[MC33_OFF_CASES + 2 * id + 1];
            row = select_tiling(q, len, need_c, stale);
            if (need_c) { compute_c_vertex(q); atomicMax(&s_maxw, j); atomicOr(&s_wmask[j >> 5], 1u << (j & 31)); s_v12[3 * j] = q.v12.x; s_v12[3 * j + 1] = q.v12.y; s_v12[3 * j + 2] = q.v12.z; }
        }
        const int any_stale = __syncthreads_or(stale);
        bool unres = false;
        if (any_stale && stale) {  // nearest earlier cell of this brick that computed a c-vertex, else the carry entering the brick
            int widx = -1;
            const int wd = j >> 5, bit = j & 31;
            const unsigned m = s_wmask[wd] & ((1u << bit) - 1u);
            if (m) widx = wd * 32 + (31 - __clz(m));
            else for (int k = wd - 1; k >= 0; --k) if (s_wmask[k]) { widx = k * 32 + (31 - __clz(s_wmask[k])); break; }
            if (widx >= 0) q.v12 = f3{s_v12[3 * widx], s_v12[3 * widx + 1], s_v12[3 * widx + 2]};
            else if (incoming) { const CarryV12 cv = incoming[b]; if (cv.valid) q.v12 = f3{cv.x, cv.y, cv.z}; }
            else unres = true;
        }
        if (any_stale && !WRITE && !only_flagged) brick_unres = brick_unres || (__syncthreads_or(unres) != 0);
        int n = 0;
        if (len) n = emit_rows<false>(q, vs, nullptr, row, len);
        int excl, total;
        Scan(tmp).ExclusiveSum(n, excl, total);
        if (WRITE && n) emit_rows<true>(q, vs, out + (item_offsets[item] + running + (unsigned long long)excl) * 9, row, len);
        running += (unsigned long long)total;
        __syncthreads();
    }
    if (!WRITE && tid == 0) {
        item_counts[item] = (unsigned)running;
        if (!only_flagged) {
            const int mw = s_maxw;
            writer[b] = mw >= 0 ? CarryV12{s_v12[3 * mw], s_v12[3 * mw + 1], s_v12[3 * mw + 2], 1} : CarryV12{0.f, 0.f, 0.f, 0};
            unresolved[b] = (unsigned char)brick_unres;
            if (brick_unres) *any_unresolved = 1;
        }
    }
}

// ---- extraction of volumes without active tiles: one WARP per brick ----------------------------------------------------------
// The warp stages the 9^3 values, builds 9-bit active / sign rows per (x, y) from the mask bytes and the staged values,
// classifies all 512 cells with a handful of bitwise operations per column and runs the MC33 logic on the compacted candidates
// 32 at a time. The reference's stale c-vertex (tiling 6.1.2) travels from brick to brick through a chain of per-brick
// descriptors (decoupled look-back). No block-wide barriers anywhere.
constexpr int MCF_WARPS = 4;
constexpr unsigned FULL = 0xFFFFFFFFu;
struct McDesc {
    unsigned long long* count;  // [n] state << 62 | triangles: state 1 = this brick only, 2 = all bricks up to and including this one
    float* carry;               // [n][4] x, y, z, state: 1 = brick computes no c-vertex (look further back), 2 = xyz is the c-vertex leaving the brick
    unsigned* ticket;
    const float* init;          // c-vertex entering the volume (x, y, z, valid) when the volume continues an earlier slab (bs_voxel_remesh_into); else null
};
__device__ __forceinline__ float ld_volf(const float* p) { return *(const volatile float*)p; }

// c-vertex entering brick `tile`: nearest earlier brick that computed one, else the reference's initial (0, 0, 0)
__device__ f3 mcf_incoming_carry(const McDesc& D, long long tile, unsigned lane) {
    for (long long base = tile - 1; base >= 0; base -= 32) {
        const long long idx = base - lane;
        float st = 1.f;
        if (idx >= 0) { while ((st = ld_volf(D.carry + 4 * idx + 3)) == 0.f) __nanosleep(100); }
        const unsigned m = __ballot_sync(FULL, st == 2.f);
        if (m) {
            const int src = __ffs(m) - 1;  // lane 0 looks at the nearest predecessor
            f3 v{0.f, 0.f, 0.f};
            if ((int)lane == src) { __threadfence(); v = f3{ld_volf(D.carry + 4 * idx), ld_volf(D.carry + 4 * idx + 1), ld_volf(D.carry + 4 * idx + 2)}; }
            return f3{__shfl_sync(FULL, v.x, src), __shfl_sync(FULL, v.y, src), __shfl_sync(FULL, v.z, src)};
        }
    }
    if (D.init && D.init[3] != 0.f) return f3{D.init[0], D.init[1], D.init[2]};  // left behind by the previous slab
    return f3{0.f, 0.f, 0.f};
}
// c-vertex leaving the volume (all descriptors final): the next slab of a pipelined remesh starts from it. Most bricks compute
// none, so the last one that does is found by all bricks in parallel (slot = D.ticket + 1, zeroed with the ticket, holds index + 1).
__global__ void k_mc_last_carry(McDesc D, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n && D.carry[4 * i + 3] == 2.f) atomicMax(D.ticket + 1, (unsigned)(i + 1));
}
__global__ void k_mc_final_carry(McDesc D, float* out /*x, y, z, valid*/) {
    const unsigned last = D.ticket[1];
    if (last) { const long long i = (long long)last - 1; out[0] = D.carry[4 * i]; out[1] = D.carry[4 * i + 1]; out[2] = D.carry[4 * i + 2]; out[3] = 1.f; }
    else if (!(D.init && D.init[3] != 0.f)) { out[0] = out[1] = out[2] = out[3] = 0.f; }  // (otherwise the incoming one passes through: out == init)
}

struct McfCarry { f3 run; bool have_run; f3 incoming; bool incoming_known; };
// Give a 6.1.2 cell the c-vertex the reference would hold at that point: latest earlier cell of this round / of earlier
// rounds of the brick / of earlier bricks. Then advance the brick's running c-vertex. Warp-uniform control flow.
__device__ __forceinline__ void mcf_resolve(Cell& q, bool need_c, bool stale, McfCarry& C, const McDesc& D, long long tile, unsigned lane) {
    const unsigned wmask = __ballot_sync(FULL, need_c);
    const unsigned smask = __ballot_sync(FULL, stale);
    if (smask) {
        const unsigned m = wmask & ((1u << lane) - 1u);
        const int src = m ? 31 - __clz(m) : (int)lane;
        const float sx = __shfl_sync(FULL, q.v12.x, src), sy = __shfl_sync(FULL, q.v12.y, src), sz = __shfl_sync(FULL, q.v12.z, src);
        const bool outside = stale && !m && !C.have_run;
        if (__any_sync(FULL, outside) && !C.incoming_known) { C.incoming = mcf_incoming_carry(D, tile, lane); C.incoming_known = true; }
        if (stale) q.v12 = m ? f3{sx, sy, sz} : (C.have_run ? C.run : C.incoming);
    }
    if (wmask) {
        const int top = 31 - __clz(wmask);
        const float rx = __shfl_sync(FULL, q.v12.x, top), ry = __shfl_sync(FULL, q.v12.y, top), rz = __shfl_sync(FULL, q.v12.z, top);
        C.run = f3{rx, ry, rz}; C.have_run = true;
    }
}
__device__ __forceinline__ void mcf_load_cell(Cell& q, const float* val, int ox, int oy, int oz, unsigned c, int& id) {
    const unsigned x = c >> 6, y = (c >> 3) & 7, z = c & 7;
    id = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float v = val[(x + c_corner[i][0]) * 81 + (y + c_corner[i][1]) * 9 + (z + c_corner[i][2])];
        if (fabsf(v) < MIN_ABS) v = copysignf(MIN_ABS, v);
        if (v < 0.f) id |= 1 << i;
        q.c[i] = v;
    }
    q.ox = ox + (int)x; q.oy = oy + (int)y; q.oz = oz + (int)z;
}

// count (ticketed, carries the c-vertex chain) -> exclusive scan -> emit. k_mc_count does stage + classify + tiling + exact
// triangle count per brick (no ordering between bricks except the rare c-vertex look-back), a device scan turns the counts
// into offsets, and k_mc_emit -- embarrassingly parallel, bricks without output skipped before anything is staged --
// recomputes the candidates, stages each round's triangles in shared memory and copies them out with coalesced stores, to
// one destination or, on a sharded volume, straight into every rank's result buffer over NVLink (the output exchange of the
// multi-GPU path rides on the emission instead of following it). Round 1's one-pass kernel (per-brick output offsets from a
// decoupled look-back, 9 scattered floats stored per lane and triangle) took 3.3 ms at config 5: its warps waited in the
// look-back about half the time (bricks retire in order); count 1.15 ms + emit 1.45 ms now.
struct McStageS {
    float val[732];
    unsigned short rowA[82], rowS[82];  // per (x, y), x, y in 0..8: bit z = active / negative
    unsigned short cells[512];          // candidate cells in cell order
};
constexpr int MC_STG = 128;  // triangles a round can stage (32 candidates x ~2.4 triangles on average; more: the round goes in two halves, then unstaged)
struct McEmitS { McStageS g; float tri[MC_STG * 9]; };

// stage brick b (+ halo) into s, classify its 512 cells; returns the number of candidate cells (listed in s.cells)
__device__ __forceinline__ int mcs_stage_classify(const VolView& V, size_t b, McStageS& s, unsigned lane, int& ox, int& oy, int& oz) {
    int mynb = -1;
    if (lane < 8) mynb = V.nbr[b * 8 + lane];
    { int bx, by, bz; bs_key_brick(V.keys[b], bx, by, bz); ox = bx << 3; oy = by << 3; oz = bz << 3; }
    {
        const float4* g4 = reinterpret_cast<const float4*>(V.values + b * 512);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned q4 = lane + 32 * k, off = q4 * 4;
            const float4 v = g4[q4];
            float* d = s.val + (off >> 6) * 81 + ((off >> 3) & 7) * 9 + (off & 7);
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
    }
#pragma unroll
    for (int it = 0; it < 7; ++it) {  // the 217 halo entries: x = 8 plane (81), y = 8 plane without x = 8 (72), z = 8 plane without x, y = 8 (64)
        const unsigned i = lane + 32 * it;
        unsigned x = 8, y = 0, z = 0;
        if (i < 81) { y = i / 9; z = i % 9; }
        else if (i < 153) { x = (i - 81) / 9; y = 8; z = (i - 81) % 9; }
        else { x = (i - 153) >> 3; y = (i - 153) & 7; z = 8; }
        const int src = __shfl_sync(FULL, mynb, (x >> 3) | ((y >> 3) << 1) | ((z >> 3) << 2));
        if (i < 217) s.val[x * 81 + y * 9 + z] = src >= 0 ? V.values[(size_t)src * 512 + (((x & 7) << 6) | ((y & 7) << 3) | (z & 7))] : 0.f;
    }
    const unsigned char* mbytes = reinterpret_cast<const unsigned char*>(V.masks);  // byte y of word x = the z-row (x, y)
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const unsigned r = lane + 32 * it;
        const unsigned x = r < 81 ? r / 9 : 0, y = r < 81 ? r % 9 : 0;
        const unsigned nbi = (x >> 3) | ((y >> 3) << 1);
        const int s0 = __shfl_sync(FULL, mynb, nbi), s1 = __shfl_sync(FULL, mynb, nbi | 4);
        if (r < 81) {
            const unsigned bo = (x & 7) * 8 + (y & 7);
            unsigned a = s0 >= 0 ? mbytes[(size_t)s0 * 64 + bo] : 0u;
            if (s1 >= 0) a |= (unsigned)(mbytes[(size_t)s1 * 64 + bo] & 1u) << 8;
            s.rowA[r] = (unsigned short)a;
        }
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const unsigned r = lane + 32 * it;
        if (r < 81) {
            const float* p = s.val + (r / 9) * 81 + (r % 9) * 9;
            unsigned m = 0;
#pragma unroll
            for (int z = 0; z < 9; ++z) m |= (__float_as_uint(p[z]) >> 31) << z;
            s.rowS[r] = (unsigned short)m;
        }
    }
    __syncwarp();
    // a cell is a candidate if its 8 corners are active and their signs differ
    unsigned cross[2], cnt[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const unsigned c = lane + 32 * h, x = c >> 3, y = c & 7, r = x * 9 + y;
        const unsigned a = s.rowA[r] & s.rowA[r + 1] & s.rowA[r + 9] & s.rowA[r + 10];
        const unsigned sa = s.rowS[r] & s.rowS[r + 1] & s.rowS[r + 9] & s.rowS[r + 10];
        const unsigned so = s.rowS[r] | s.rowS[r + 1] | s.rowS[r + 9] | s.rowS[r + 10];
        cross[h] = (a & (a >> 1)) & (so | (so >> 1)) & ~(sa & (sa >> 1)) & 0xFFu;
        cnt[h] = __popc(cross[h]);
    }
    unsigned inc0 = cnt[0], inc1 = cnt[1];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t0 = __shfl_up_sync(FULL, inc0, o), t1 = __shfl_up_sync(FULL, inc1, o);
        if ((int)lane >= o) { inc0 += t0; inc1 += t1; }
    }
    const unsigned tot0 = __shfl_sync(FULL, inc0, 31), tot1 = __shfl_sync(FULL, inc1, 31);
    {
        unsigned o0 = inc0 - cnt[0], o1 = tot0 + inc1 - cnt[1];
        for (unsigned m = cross[0]; m; m &= m - 1) s.cells[o0++] = (unsigned short)((lane << 3) | (__ffs(m) - 1));
        for (unsigned m = cross[1]; m; m &= m - 1) s.cells[o1++] = (unsigned short)(((lane + 32) << 3) | (__ffs(m) - 1));
    }
    __syncwarp();
    return (int)(tot0 + tot1);
}

__global__ void __launch_bounds__(32 * MCF_WARPS, 8) k_mc_count(VolView V, const signed char* __restrict__ tables, float vs, McDesc D, unsigned* counts) {
    __shared__ McStageS S[MCF_WARPS];
    const unsigned lane = threadIdx.x & 31;
    McStageS& s = S[threadIdx.x >> 5];
    for (;;) {
        unsigned tk = 0;
        if (lane == 0) tk = atomicAdd(D.ticket, 1u);  // ticket order: a brick's predecessors are running or done, so the c-vertex look-back cannot deadlock
        tk = __shfl_sync(FULL, tk, 0);
        if (tk >= V.n) break;
        const long long tile = tk;
        const size_t b = tk;
        if (V.owned && !V.owned[b]) {  // halo brick of a sharded volume: emits nothing, passes the c-vertex through
            if (lane == 0) { *(volatile float*)(D.carry + 4 * tile + 3) = 1.f; counts[b] = 0u; }
            continue;
        }
        int ox, oy, oz;
        const int n_act = mcs_stage_classify(V, b, s, lane, ox, oy, oz);
        McfCarry C{f3{0.f, 0.f, 0.f}, false, f3{0.f, 0.f, 0.f}, false};
        unsigned total = 0;
        for (int base = 0; base < n_act; base += 32) {
            const int j = base + (int)lane;
            Cell q; q.T = tables; q.v12 = f3{0.f, 0.f, 0.f};
            int row = 0, len = 0; bool need_c = false, stale = false;
            if (j < n_act) {
                int id;
                mcf_load_cell(q, s.val, ox, oy, oz, s.cells[j], id);
                q.cs = tables[MC33_OFF_CASES + 2 * id]; q.cf = tables[MC33_OFF_CASES + 2 * id + 1];
                row = select_tiling(q, len, need_c, stale);
                if (need_c) compute_c_vertex(q);
            }
            mcf_resolve(q, need_c, stale, C, D, tile, lane);
            int n = 0;
            if (len) n = emit_rows<false>(q, vs, nullptr, row, len);  // exact: degenerate triangles are dropped like the reference does
#pragma unroll
            for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(FULL, n, o);
            total += (unsigned)n;
        }
        if (lane == 0) {
            float* cd = D.carry + 4 * tile;
            const bool val = C.have_run || C.incoming_known;
            if (val) { const f3 v = C.have_run ? C.run : C.incoming; *(volatile float*)(cd) = v.x; *(volatile float*)(cd + 1) = v.y; *(volatile float*)(cd + 2) = v.z; __threadfence(); }
            *(volatile float*)(cd + 3) = val ? 2.f : 1.f;
            counts[b] = total;
        }
        __syncwarp();
    }
}

struct McDst { float* p[16]; int world; };  // where the triangles go: one buffer, or the same place in every rank's buffer (peer memory)
__global__ void __launch_bounds__(32 * MCF_WARPS, 6) k_mc_emit(VolView V, const signed char* __restrict__ tables, float vs, McDesc D, const unsigned* __restrict__ counts,
                                                                const unsigned long long* __restrict__ offsets, McDst O, unsigned long long cap_tris) {
    __shared__ McEmitS S[MCF_WARPS];
    const unsigned lane = threadIdx.x & 31;
    McEmitS& se = S[threadIdx.x >> 5];
    McStageS& s = se.g;
    // persistent warps draw bricks from a ticket counter (order does not matter here: every brick knows its offset), so a
    // warp whose brick has no output -- a third of them at config 5 -- is not idle until its CTA retires
    for (;;) {
    unsigned tk = 0;
    if (lane == 0) tk = atomicAdd(D.ticket + 2, 1u);
    tk = __shfl_sync(FULL, tk, 0);
    if (tk >= V.n) break;
    const size_t b = tk;
    const unsigned total = counts[b];
    if (total == 0) continue;  // (halo bricks, bricks without a sign change: nothing is staged)
    unsigned long long running = offsets[b];
    if (running + total > cap_tris) continue;  // the host grows the buffer and runs the emit pass again
    int ox, oy, oz;
    const int n_act = mcs_stage_classify(V, b, s, lane, ox, oy, oz);
    const long long tile = (long long)b;
    McfCarry C{f3{0.f, 0.f, 0.f}, false, f3{0.f, 0.f, 0.f}, false};  // (the carry chain is final: look-backs do not wait)
    for (int base = 0; base < n_act; base += 32) {
        const int j = base + (int)lane;
        Cell q; q.T = tables; q.v12 = f3{0.f, 0.f, 0.f};
        int row = 0, len = 0; bool need_c = false, stale = false;
        if (j < n_act) {
            int id;
            mcf_load_cell(q, s.val, ox, oy, oz, s.cells[j], id);
            q.cs = tables[MC33_OFF_CASES + 2 * id]; q.cf = tables[MC33_OFF_CASES + 2 * id + 1];
            row = select_tiling(q, len, need_c, stale);
            if (need_c) compute_c_vertex(q);
        }
        mcf_resolve(q, need_c, stale, C, D, tile, lane);
        const int nmax_all = len / 3;  // triangles of the tiling row; fewer are emitted only when one is degenerate
        int inc_all = nmax_all;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc_all, o); if ((int)lane >= o) inc_all += t; }
        const int halves = __shfl_sync(FULL, inc_all, 31) <= MC_STG ? 1 : 2;
        for (int h = 0; h < halves; ++h) {
            const bool in = halves == 1 || (int)(lane >> 4) == h;
            const int nmax = in ? nmax_all : 0;
            int incm = nmax;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, incm, o); if ((int)lane >= o) incm += t; }
            const int totm = __shfl_sync(FULL, incm, 31);
            bool fast = totm <= MC_STG;
            int n = 0;
            if (fast) {
                if (nmax) n = emit_rows<true>(q, vs, se.tri + (incm - nmax) * 9, row, len);
                fast = !__any_sync(FULL, n != nmax);
            }
            if (fast) {  // the staged triangles are dense: coalesced copy to their final place
                __syncwarp();
                const size_t at = (size_t)running * 9;
#pragma unroll 1
                for (int d = 0; d < O.world; ++d) {  // (not unrolled: 16 copies of this loop and of emit_rows below thrashed the instruction cache)
                    float* dst = O.p[d] + at;
                    for (int i = (int)lane; i < totm * 9; i += 32) dst[i] = se.tri[i];
                }
                running += (unsigned long long)totm;
                __syncwarp();
            } else {  // a degenerate triangle was dropped (or the half does not fit): exact counts, then direct stores
                n = nmax ? emit_rows<false>(q, vs, nullptr, row, len) : 0;
                int inc = n;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if ((int)lane >= o) inc += t; }
#pragma unroll 1
                for (int d = 0; d < O.world; ++d) if (n) emit_rows<true>(q, vs, O.p[d] + (running + (unsigned long long)(inc - n)) * 9, row, len);
                running += (unsigned long long)__shfl_sync(FULL, inc, 31);
                __syncwarp();
            }
        }
    }
    __syncwarp();
    }
}
struct WidenU32 { __host__ __device__ unsigned long long operator()(unsigned v) const { return v; } };

__global__ void k_shift_carry(const CarryV12* __restrict__ incl, CarryV12* excl, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) excl[i] = i ? incl[i - 1] : CarryV12{0.f, 0.f, 0.f, 0};
}

__global__ void k_mc_neighbours(const unsigned long long* __restrict__ keys, size_t n, int* nbr) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n * 8) return;
    const size_t b = i >> 3; const unsigned d = (unsigned)(i & 7);
    if (d == 0) { nbr[i] = (int)b; return; }
    int bx, by, bz; bs_key_brick(keys[b], bx, by, bz);
    bx += d & 1; by += (d >> 1) & 1; bz += d >> 2;
    nbr[i] = (bx <= BS_BRICK_MAX && by <= BS_BRICK_MAX && bz <= BS_BRICK_MAX) ? (int)find_key(keys, n, bs_brick_key(bx, by, bz)) : -1;
}

// TreeNode::at on the flat volume: a brick voxel if active, else the value of an active tile covering it
__device__ bool value_at(const VolView& V, int x, int y, int z, float& v) {
    const unsigned long long k = bs_brick_key(x >> 3, y >> 3, z >> 3);
    long long i = find_key(V.keys, V.n, k);
    if (i >= 0) {
        const unsigned off = ((x & 7) << 6) | ((y & 7) << 3) | (z & 7);
        if (!((V.masks[i * 8 + (off >> 6)] >> (off & 63)) & 1)) return false;
        v = V.values[i * 512 + off];
        return true;
    }
    i = V.nt8 ? find_key(V.t8k, V.nt8, k) : -1;
    if (i >= 0) { v = V.t8v[i]; return true; }
    i = V.nt128 ? find_key(V.t128k, V.nt128, k >> 12) : -1;
    if (i >= 0) { v = V.t128v[i]; return true; }
    return false;
}

// Active tiles (CubesVisitor::tile, marching_cubes.rs:970-994): only the boundary voxels of the tile are tested, in
// the reference's order -- for i, for j: left, right, top, bottom, front, back -- so edge and corner voxels are
// visited (and their triangles emitted) more than once, exactly like the reference. One CTA per tile.
template <bool WRITE>
__global__ void __launch_bounds__(256) k_mc_tiles(VolView V, const signed char* __restrict__ tables, float vs, int level /*0: 8^3, 1: 128^3*/,
                                                  const unsigned* __restrict__ pos, unsigned* item_counts, const unsigned long long* __restrict__ item_offsets, float* out) {
    const size_t ti = blockIdx.x;
    const size_t item = pos[ti];
    int bx, by, bz;
    if (level == 0) bs_key_brick(V.t8k[ti], bx, by, bz); else bs_key_brick(V.t128k[ti] << 12, bx, by, bz);
    const int s = level == 0 ? 8 : 128;
    const int ox = bx << 3, oy = by << 3, oz = bz << 3;
    const unsigned n_visits = 6u * s * s;
    typedef cub::BlockScan<int, 256> Scan;
    __shared__ typename Scan::TempStorage tmp;
    unsigned long long running = 0;
    for (unsigned base = 0; base < n_visits; base += 256) {
        const unsigned vi = base + threadIdx.x;
        int n = 0;
        float local[WRITE ? 12 * 9 : 1];
        if (vi < n_visits) {
            const int face = vi % 6, ij = vi / 6, i = ij / s, j = ij % s;
            int x, y, z;
            switch (face) {
                case 0: x = 0; y = i; z = j; break;          // left
                case 1: x = s - 1; y = i; z = j; break;      // right
                case 2: x = i; y = j; z = s - 1; break;      // top
                case 3: x = i; y = j; z = 0; break;          // bottom
                case 4: x = i; y = s - 1; z = j; break;      // front
                default: x = i; y = 0; z = j; break;         // back
            }
            Cell q; q.T = tables; q.v12 = f3{0.f, 0.f, 0.f};
            q.ox = ox + x; q.oy = oy + y; q.oz = oz + z;
            int id = 0; bool all = true;
            for (int c = 0; c < 8 && all; ++c) {
                float v;
                all = value_at(V, q.ox + c_corner[c][0], q.oy + c_corner[c][1], q.oz + c_corner[c][2], v);
                if (!all) break;
                if (fabsf(v) < MIN_ABS) v = copysignf(MIN_ABS, v);
                if (v < 0.f) id |= 1 << c;
                q.c[c] = v;
            }
            if (all && id != 0 && id != 255) {
                q.cs = tables[MC33_OFF_CASES + 2 * id]; q.cf = tables[MC33_OFF_CASES + 2 * id + 1];
                n = emit_cell<WRITE>(q, vs, local);
            }
        }
        int excl, total;
        Scan(tmp).ExclusiveSum(n, excl, total);
        __syncthreads();
        if (WRITE) {
            float* dst = out + (item_offsets[item] + running + (unsigned long long)excl) * 9;
            for (int k = 0; k < n * 9; ++k) dst[k] = local[k];
        }
        running += (unsigned long long)total;
    }
    if (!WRITE && threadIdx.x == 0) item_counts[item] = (unsigned)running;
}

// rank of every brick / tile in the merged key order (a 128^3 tile sorts by its 16^3-node key)
__global__ void k_merge_pos(VolView V, unsigned* pos_brick, unsigned* pos_t8, unsigned* pos_t128) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    auto lb = [](const unsigned long long* keys, size_t n, unsigned long long k, int shift) {
        size_t lo = 0, hi = n;
        while (lo < hi) { size_t mid = (lo + hi) >> 1; if ((keys[mid] << shift) < k) lo = mid + 1; else hi = mid; }
        return (unsigned)lo;
    };
    if (i < V.n) pos_brick[i] = (unsigned)i + lb(V.t8k, V.nt8, V.keys[i], 0) + lb(V.t128k, V.nt128, V.keys[i], 12);
    if (i < V.nt8) pos_t8[i] = (unsigned)i + lb(V.keys, V.n, V.t8k[i], 0) + lb(V.t128k, V.nt128, V.t8k[i], 12);
    if (i < V.nt128) pos_t128[i] = (unsigned)i + lb(V.keys, V.n, V.t128k[i] << 12, 0) + lb(V.t8k, V.nt8, V.t128k[i] << 12, 0);
}

__global__ void k_widen(const unsigned* in, unsigned long long* out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
    if (i == n) out[i] = 0;
}

}  // namespace

bs_status bs_ensure_out_verts(bs_context* ctx, size_t n_floats) {
    if (ctx->out_verts_cap < n_floats) {
        if (ctx->d_out_verts) cudaFree(ctx->d_out_verts);
        ctx->d_out_verts = nullptr; ctx->out_verts_cap = 0;
        size_t cap = n_floats + n_floats / 4 + 1024;
        BS_CUDA(ctx, cudaMalloc((void**)&ctx->d_out_verts, cap * sizeof(float)));
        ctx->out_verts_cap = cap;
    }
    return BS_OK;
}

// ---- marching cubes in two steps (volumes without active tiles): the count leaves its per-brick products in the context, the emit
// consumes them. bs_mc_impl runs both; the multi-GPU path runs them as two ABI calls with the ranks' count exchange in between.
void bs_mc_pending_release(bs_context* ctx) {
    bs_mc_pending& M = ctx->mc_pending;
    bs_free(ctx, M.d_desc); bs_free(ctx, M.d_counts); bs_free(ctx, M.d_offsets); bs_free(ctx, M.d_nbr);
    M = bs_mc_pending();
}
bs_status bs_mc_count_phase(const bs_volume* v, float voxel_size, size_t* n_verts) {
    bs_context* ctx = v->ctx;
    cudaStream_t st = ctx->stream;
    bs_mc_pending& M = ctx->mc_pending;
    const size_t n = v->n_bricks;
    *n_verts = 0;
    M.vol = v; M.voxel_size = voxel_size; M.n = n; M.n_tris = 0;
    if (n == 0) return BS_OK;
    BS_TRY(bs_alloc(ctx, &M.d_nbr, n * 8));
    bs_count_launch(), k_mc_neighbours<<<bs_blocks(n * 8, 256), 256, 0, st>>>(v->keys, n, M.d_nbr);
    VolView V{v->keys, v->values, v->masks, n, v->owned, M.d_nbr, nullptr, nullptr, 0, nullptr, nullptr, 0};
    BS_TRY(bs_alloc(ctx, &M.d_desc, 2 * n + 2)); BS_TRY(bs_alloc(ctx, &M.d_counts, n + 1)); BS_TRY(bs_alloc(ctx, &M.d_offsets, n + 1));
    McDesc D{nullptr, reinterpret_cast<float*>(M.d_desc), reinterpret_cast<unsigned*>(M.d_desc + 2 * n), ctx->mc_chain ? ctx->d_mc_carry : nullptr};
    static bool carve = false;
    if (!carve) {
        cudaFuncSetAttribute(k_mc_count, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_mc_emit, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        carve = true;
    }
    BS_CUDA(ctx, cudaMemsetAsync(M.d_desc, 0, (2 * n + 2) * sizeof(unsigned long long), st));  // carry records, count ticket, last-carry slot, emit ticket
    BS_CUDA(ctx, cudaMemsetAsync(M.d_counts + n, 0, sizeof(unsigned), st));
    const unsigned grid0 = (unsigned)std::min<size_t>((n + MCF_WARPS - 1) / MCF_WARPS, (size_t)ctx->sm_count * 8);
    bs_count_launch(), k_mc_count<<<grid0, 32 * MCF_WARPS, 0, st>>>(V, (const signed char*)ctx->d_mc33, voxel_size, D, M.d_counts);
    void* d_tmp = nullptr; size_t tmp_bytes = 0;
    auto wide = thrust::make_transform_iterator((const unsigned*)M.d_counts, WidenU32());
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, wide, M.d_offsets, (int)(n + 1), st);
    BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp_bytes));
    cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, wide, M.d_offsets, (int)(n + 1), st);
    bs_free(ctx, d_tmp);
    if (ctx->mc_chain) { bs_count_launch(), k_mc_last_carry<<<bs_blocks(n, 256), 256, 0, st>>>(D, (long long)n); bs_count_launch(), k_mc_final_carry<<<1, 1, 0, st>>>(D, ctx->d_mc_carry); }
    bs_mark(ctx, "mc_count_ms");
    BS_TRY(bs_fetch(ctx, &M.n_tris, M.d_offsets + n, sizeof(M.n_tris)));
    BS_TRY(bs_sync(ctx));
    *n_verts = (size_t)M.n_tris * 3;
    return BS_OK;
}
bs_status bs_mc_emit_phase(const bs_volume* v, float* const* dst, int world, size_t offset_floats, size_t cap_floats) {
    bs_context* ctx = v->ctx;
    cudaStream_t st = ctx->stream;
    bs_mc_pending& M = ctx->mc_pending;
    if (M.vol != v) return bs_fail(ctx, BS_ERR_INVALID, "marching cubes emit without a count on this volume");
    const size_t n = M.n;
    if (n && M.n_tris) {
        if (offset_floats + (size_t)M.n_tris * 9 > cap_floats) return bs_fail(ctx, BS_ERR_INVALID, "marching cubes: destination too small");
        VolView V{v->keys, v->values, v->masks, n, v->owned, M.d_nbr, nullptr, nullptr, 0, nullptr, nullptr, 0};
        McDesc D{nullptr, reinterpret_cast<float*>(M.d_desc), reinterpret_cast<unsigned*>(M.d_desc + 2 * n), ctx->mc_chain ? ctx->d_mc_carry : nullptr};
        McDst O; O.world = world;
        for (int d = 0; d < 16; ++d) O.p[d] = d < world ? dst[d] + offset_floats : nullptr;
        const unsigned grid = (unsigned)std::min<size_t>((n + MCF_WARPS - 1) / MCF_WARPS, (size_t)ctx->sm_count * 6);
        bs_count_launch(), k_mc_emit<<<grid, 32 * MCF_WARPS, 0, st>>>(V, (const signed char*)ctx->d_mc33, M.voxel_size, D, M.d_counts, M.d_offsets, O, ~0ull);
    }
    bs_mark(ctx, "mc_emit_ms");
    return BS_OK;
}

bs_status bs_mc_impl(const bs_volume* v, float voxel_size, const float** d_verts, size_t* n_verts) {
    bs_context* ctx = v->ctx;
    cudaStream_t st = ctx->stream;
    *d_verts = nullptr; *n_verts = 0;
    bs_marks_begin(ctx);
    const size_t n = v->n_bricks, nt8 = v->n_tiles8, nt128 = v->n_tiles128, n_items = n + nt8 + nt128;
    if (n_items == 0) { bs_marks_end(ctx); return BS_OK; }
    if (!(nt8 || nt128)) {
        // count (ticketed) -> scan -> emit into the context's result buffer (grown to the exact size when it is too small)
        bs_mc_pending_release(ctx);
        size_t nv = 0;
        BS_TRY(bs_mc_count_phase(v, voxel_size, &nv));
        const unsigned long long n_tris = nv / 3;
        bs_status s = BS_OK;
        if (ctx->out_verts_cap == 0) s = bs_ensure_out_verts(ctx, n * 160 * 9);
        if (s == BS_OK && (size_t)n_tris * 9 > ctx->out_verts_cap) s = bs_ensure_out_verts(ctx, (size_t)n_tris * 9);
        if (s == BS_OK) { float* one[1] = {ctx->d_out_verts}; s = bs_mc_emit_phase(v, one, 1, 0, ctx->out_verts_cap); }
        bs_mc_pending_release(ctx);
        if (s != BS_OK) return s;
        BS_CUDA(ctx, cudaGetLastError());
        bs_marks_end(ctx);
        bs_stat_add(ctx, "n_bricks", (double)n);
        bs_stat_add(ctx, "n_out_tris", (double)n_tris);
        *d_verts = ctx->d_out_verts; *n_verts = (size_t)n_tris * 3;
        return BS_OK;
    }
    int* d_nbr = nullptr;
    BS_TRY(bs_alloc(ctx, &d_nbr, n * 8));
    if (n) bs_count_launch(), k_mc_neighbours<<<bs_blocks(n * 8, 256), 256, 0, st>>>(v->keys, n, d_nbr);
    VolView V{v->keys, v->values, v->masks, n, v->owned, d_nbr, v->tile8_keys, v->tile8_values, nt8, v->tile128_keys, v->tile128_values, nt128};
    const signed char* tables = (const signed char*)ctx->d_mc33;
    // volumes with active tiles (CSG unions): count pass, scan over bricks and tiles in merged order, emit pass
    unsigned *d_counts = nullptr, *d_pos = nullptr; unsigned long long *d_wide = nullptr, *d_off = nullptr;
    BS_TRY(bs_alloc(ctx, &d_counts, n_items)); BS_TRY(bs_alloc(ctx, &d_wide, n_items + 1)); BS_TRY(bs_alloc(ctx, &d_off, n_items + 1));
    const bool tiles = nt8 || nt128;
    if (tiles) {
        BS_TRY(bs_alloc(ctx, &d_pos, n_items));
        bs_count_launch(), k_merge_pos<<<bs_blocks(std::max(n, std::max(nt8, nt128)), 256), 256, 0, st>>>(V, d_pos, d_pos + n, d_pos + n + nt8);
    }
    CarryV12 *d_writer = nullptr, *d_carry_incl = nullptr, *d_incoming = nullptr; unsigned char* d_unres = nullptr; int* d_any = nullptr; int any_unres = 0;
    BS_TRY(bs_alloc(ctx, &d_writer, n)); BS_TRY(bs_alloc(ctx, &d_unres, n)); BS_TRY(bs_alloc(ctx, &d_any, 1));
    BS_CUDA(ctx, cudaMemsetAsync(d_any, 0, sizeof(int), st));
    if (n) bs_count_launch(), k_mc<false><<<(unsigned)n, MC_TPB, 0, st>>>(V, tables, voxel_size, d_pos, d_counts, nullptr, nullptr, d_writer, nullptr, d_unres, 0, d_any);
    BS_TRY(bs_fetch(ctx, &any_unres, d_any, sizeof(int)));
    BS_TRY(bs_sync(ctx));
    if (any_unres) {  // rare: some 6.1.2 cell needs the c-vertex left behind by an earlier brick -> carry scan, recount those bricks
        BS_TRY(bs_alloc(ctx, &d_carry_incl, n)); BS_TRY(bs_alloc(ctx, &d_incoming, n));
        void* d_t2 = nullptr; size_t t2 = 0;
        cub::DeviceScan::InclusiveScan(nullptr, t2, d_writer, d_carry_incl, CarryOp(), n, st);
        BS_TRY(bs_alloc(ctx, (char**)&d_t2, t2));
        cub::DeviceScan::InclusiveScan(d_t2, t2, d_writer, d_carry_incl, CarryOp(), n, st);
        bs_count_launch(), k_shift_carry<<<bs_blocks(n, 256), 256, 0, st>>>(d_carry_incl, d_incoming, n);
        bs_count_launch(), k_mc<false><<<(unsigned)n, MC_TPB, 0, st>>>(V, tables, voxel_size, d_pos, d_counts, nullptr, nullptr, d_writer, d_incoming, d_unres, 1, d_any);
        bs_free(ctx, d_t2);
    }
    if (nt8) bs_count_launch(), k_mc_tiles<false><<<(unsigned)nt8, 256, 0, st>>>(V, tables, voxel_size, 0, d_pos + n, d_counts, nullptr, nullptr);
    if (nt128) bs_count_launch(), k_mc_tiles<false><<<(unsigned)nt128, 256, 0, st>>>(V, tables, voxel_size, 1, d_pos + n + nt8, d_counts, nullptr, nullptr);
    bs_count_launch(), k_widen<<<bs_blocks(n_items + 1, 256), 256, 0, st>>>(d_counts, d_wide, n_items);
    void* d_tmp = nullptr; size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_wide, d_off, n_items + 1, st);
    BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp_bytes));
    cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_wide, d_off, n_items + 1, st);
    unsigned long long n_tris = 0;
    BS_TRY(bs_fetch(ctx, &n_tris, d_off + n_items, sizeof(n_tris)));
    BS_TRY(bs_sync(ctx));
    bs_mark(ctx, "mc_count_ms");
    bs_status s = bs_ensure_out_verts(ctx, (size_t)n_tris * 9);
    if (s == BS_OK && n_tris) {
        if (n) bs_count_launch(), k_mc<true><<<(unsigned)n, MC_TPB, 0, st>>>(V, tables, voxel_size, d_pos, nullptr, d_off, ctx->d_out_verts, nullptr, d_incoming, nullptr, 0, nullptr);
        if (nt8) bs_count_launch(), k_mc_tiles<true><<<(unsigned)nt8, 256, 0, st>>>(V, tables, voxel_size, 0, d_pos + n, nullptr, d_off, ctx->d_out_verts);
        if (nt128) bs_count_launch(), k_mc_tiles<true><<<(unsigned)nt128, 256, 0, st>>>(V, tables, voxel_size, 1, d_pos + n + nt8, nullptr, d_off, ctx->d_out_verts);
    }
    bs_mark(ctx, "mc_emit_ms");
    bs_free(ctx, d_pos); bs_free(ctx, d_nbr); bs_free(ctx, d_writer); bs_free(ctx, d_carry_incl); bs_free(ctx, d_incoming); bs_free(ctx, d_unres); bs_free(ctx, d_any);
    bs_free(ctx, d_tmp); bs_free(ctx, d_counts); bs_free(ctx, d_wide); bs_free(ctx, d_off);
    if (s != BS_OK) return s;
    BS_CUDA(ctx, cudaGetLastError());
    bs_marks_end(ctx);
    bs_stat_add(ctx, "n_bricks", (double)n);
    bs_stat_add(ctx, "n_out_tris", (double)n_tris);
    *d_verts = ctx->d_out_verts; *n_verts = (size_t)n_tris * 3;
    return BS_OK;
}
