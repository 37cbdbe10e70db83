// Dual contouring on sorted bricks (sm_100a).
// Replaces DualContouringMesher::mesh (src/voxel/meshing/dual_contouring.rs:23-83): edge intersections + normals
// (ComputeEdgeIntersectionsVisitor :248-387), per-cell feature points (ComputeCellPointsVisitor :174-246,
// find_feature_point :429-458), quads (TriangulateVisitor :93-172), offsets tables :389-423.
//
// The reference keeps Hermite data and cell points in three + one auxiliary sparse trees behind mutexes. Here:
//   pass A  one CTA per brick stages the brick with an 11^3 halo ([-1, +9] per axis) in shared memory; the cells are
//           classified, the surface cells (all 8 corners active, mixed signs) compacted, and dense lanes re-derive the
//           Hermite samples of a cell's 12 edges (gradients by central / one-sided differences from the staged values)
//           and run the 50-iteration particle solve in registers; the feature point goes to a per-brick array
//           (512 x float3) with a validity mask. The same pass checks,
//           for every sign-change edge the brick owns, that both end points have a neighbour on every axis
//           (the reference hits unreachable!() otherwise, :340).
//   pass B  one CTA per brick, one thread per voxel: for +x/+y/+z sign-change edges fetch the four surrounding
//           cell points (own brick or the -x/-y/-z neighbour bricks), emit the quad as two triangles, scaled by
//           voxel_size, degenerate ones dropped; count -> scan -> emit keeps leaf order / voxel order / X,Y,Z.
// All arithmetic uses the non-contracting helpers, so the result is bit-identical to a sequential evaluation.
#include "bs_common.cuh"
#include <cub/cub.cuh>
#include <cfloat>

namespace {

typedef unsigned long long u64;
constexpr int H = 11;          // halo edge: local coordinate l in [0, 11) <-> voxel offset l - 1 in [-1, 9]
constexpr int HN = H * H * H;  // 1331

__constant__ signed char c_corner[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
// EDGE_OFFSETS (dual_contouring.rs:408-421): lower voxel offset + direction
__constant__ signed char c_edge[12][4] = {{0, 0, 0, 0}, {0, 0, 0, 1}, {0, 0, 0, 2}, {1, 0, 0, 1}, {1, 0, 0, 2}, {0, 0, 1, 0},
                                          {0, 0, 1, 1}, {0, 1, 0, 0}, {0, 1, 0, 2}, {1, 0, 1, 1}, {0, 1, 1, 0}, {1, 1, 0, 2}};
// CELL_OFFSETS (:389-406)
__constant__ signed char c_cell[3][4][3] = {{{0, 0, 0}, {0, 0, -1}, {0, -1, -1}, {0, -1, 0}}, {{0, 0, 0}, {-1, 0, 0}, {-1, 0, -1}, {0, 0, -1}}, {{0, -1, 0}, {-1, -1, 0}, {-1, 0, 0}, {0, 0, 0}}};

__device__ __forceinline__ long long find_key(const u64* keys, size_t n, u64 k) {
    size_t lo = 0, hi = n;
    while (lo < hi) { size_t mid = (lo + hi) >> 1; if (keys[mid] < k) lo = mid + 1; else hi = mid; }
    return (lo < n && keys[lo] == k) ? (long long)lo : -1;
}
__device__ __forceinline__ int hidx(int x, int y, int z) { return (x * H + y) * H + z; }  // local coords

struct Halo { const float* v; const unsigned char* a; };

// grad (:326-342); *bad is set when the reference would hit unreachable!()
__device__ __forceinline__ float grad(const Halo& h, int x, int y, int z, int axis, bool* bad) {
    const int st = axis == 0 ? H * H : (axis == 1 ? H : 1);
    const int i = hidx(x, y, z);
    const int c = axis == 0 ? x : (axis == 1 ? y : z);
    const bool al = c > 0 && h.a[i - st], ar = c < H - 1 && h.a[i + st];
    if (al && ar) return xmul(xsub(h.v[i + st], h.v[i - st]), 0.5f);
    if (ar) return xsub(h.v[i + st], h.v[i]);
    if (al) return xsub(h.v[i], h.v[i - st]);
    *bad = true;
    return 0.f;
}
// intersection + normal of the edge (lower voxel local (x,y,z), dir); false if no sign change (:266-307)
__device__ __noinline__ bool hermite(const Halo& h, const int* org, int x, int y, int z, int dir, f3& point, f3& normal, bool* bad) {
    const int x2 = x + (dir == 0), y2 = y + (dir == 1), z2 = z + (dir == 2);
    const float v1 = h.v[hidx(x, y, z)], v2 = h.v[hidx(x2, y2, z2)];
    if (((__float_as_uint(v1) ^ __float_as_uint(v2)) >> 31) == 0) return false;
    const float t = (v1 == v2) ? 0.5f : xdiv(v1, xsub(v1, v2));
    point = f3{(float)(org[0] + x - 1), (float)(org[1] + y - 1), (float)(org[2] + z - 1)};
    if (dir == 0) point.x = xadd(point.x, t); else if (dir == 1) point.y = xadd(point.y, t); else point.z = xadd(point.z, t);
    const float omt = xsub(1.0f, t);
    float g[3];
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) g[ax] = xadd(xmul(omt, grad(h, x, y, z, ax, bad)), xmul(t, grad(h, x2, y2, z2, ax, bad)));
    const f3 n = f3{g[0], g[1], g[2]};
    const float len = xsqrt(xnorm2(n));
    normal = f3{xdiv(n.x, len), xdiv(n.y, len), xdiv(n.z, len)};
    return true;
}

// 27-neighbourhood of every brick (index into the sorted brick list, -1 = absent), one thread per (brick, neighbour): the
// binary searches run at throughput here instead of as a dependent chain at the head of every k_dc_cells CTA
__global__ void k_dc_neighbours(const u64* __restrict__ keys, size_t n, int* nbr /*[n][27]*/) {
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g >= n * 27) return;
    const size_t b = g / 27; const unsigned t = (unsigned)(g % 27);
    if (t == 13) { nbr[g] = (int)b; return; }
    int bx, by, bz; bs_key_brick(keys[b], bx, by, bz);
    const int nx = bx + (int)(t / 9) - 1, ny = by + (int)((t / 3) % 3) - 1, nz = bz + (int)(t % 3) - 1;
    const bool ok = nx >= BS_BRICK_MIN && nx <= BS_BRICK_MAX && ny >= BS_BRICK_MIN && ny <= BS_BRICK_MAX && nz >= BS_BRICK_MIN && nz <= BS_BRICK_MAX;
    nbr[g] = ok ? (int)find_key(keys, n, bs_brick_key(nx, ny, nz)) : -1;
}

// One CTA of 128 threads per brick. Only ~10 % of a band brick's cells hold a sign change: the cells are classified first
// (4 per thread), the surface cells are compacted in shared memory, and the Hermite samples + 50-iteration solve run on
// dense lanes; other cells cost a classification and nothing else (their slot of cell_pts is never read: k_dc_quads
// tests cell_valid first).
constexpr int DC_TPB = 128;
__global__ void __launch_bounds__(DC_TPB, 6) k_dc_cells(const u64* __restrict__ keys, const float* __restrict__ values, const u64* __restrict__ masks, size_t n, const int* __restrict__ nbr,
                                                        const unsigned char* __restrict__ owned, float* cell_pts /*n*512*3*/, u64* cell_valid /*n*8*/, int* flags) {
    __shared__ float s_v[HN];
    __shared__ unsigned char s_a[HN];
    __shared__ int s_nb[27];
    __shared__ int s_org[3];
    __shared__ unsigned s_bal[16];
    __shared__ unsigned short s_list[512];
    __shared__ float s_damp[50];  // 1 - it / 50 (:447), the same for every cell
    const size_t b = blockIdx.x;
    const unsigned t = threadIdx.x;
    if (t >= 64 && t < 114) s_damp[t - 64] = xsub(1.0f, xdiv((float)(t - 64), 50.0f));
    if (t < 27) s_nb[t] = nbr[b * 27 + t];
    if (t == 32) { int bx, by, bz; bs_key_brick(keys[b], bx, by, bz); s_org[0] = bx << 3; s_org[1] = by << 3; s_org[2] = bz << 3; }
    __syncthreads();
    {   // stage the 11^3 halo: all loads of a thread are issued before the first store
        constexpr int NQ = (HN + DC_TPB - 1) / DC_TPB;
        float v[NQ]; u64 m[NQ]; unsigned sh[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const unsigned i = t + DC_TPB * q;
            v[q] = 0.f; m[q] = 0; sh[q] = 0;
            if (i < HN) {
                const int lx = i / (H * H), ly = (i / H) % H, lz = i % H;        // local; voxel offset = l - 1
                const int ox = lx - 1, oy = ly - 1, oz = lz - 1;
                const int nbx = ox < 0 ? 0 : (ox > 7 ? 2 : 1), nby = oy < 0 ? 0 : (oy > 7 ? 2 : 1), nbz = oz < 0 ? 0 : (oz > 7 ? 2 : 1);
                const int src = s_nb[(nbx * 3 + nby) * 3 + nbz];
                if (src >= 0) {
                    const unsigned off = ((ox & 7) << 6) | ((oy & 7) << 3) | (oz & 7);
                    m[q] = masks[(size_t)src * 8 + (off >> 6)]; sh[q] = off & 63;
                    v[q] = values[(size_t)src * 512 + off];
                }
            }
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const unsigned i = t + DC_TPB * q;
            if (i < HN) { s_v[i] = v[q]; s_a[i] = (unsigned char)((m[q] >> sh[q]) & 1ull); }
        }
    }
    __syncthreads();
    const Halo h{s_v, s_a};
    bool bad = false;
#pragma unroll 1  // (code size: the instruction cache, not the math, bounds this kernel when everything is unrolled)
    for (int q = 0; q < 512 / DC_TPB; ++q) {
        const unsigned c = t + DC_TPB * q;
        const int x = (c >> 6) + 1, y = ((c >> 3) & 7) + 1, z = (c & 7) + 1;  // local coords of this voxel / cell origin
        // (b) every sign-change edge owned by this brick must have computable normals (else the reference panics)
        if (s_a[hidx(x, y, z)]) {
#pragma unroll 1
            for (int dir = 0; dir < 3; ++dir) {
                const int x2 = x + (dir == 0), y2 = y + (dir == 1), z2 = z + (dir == 2);
                if (!s_a[hidx(x2, y2, z2)]) continue;
                if (((__float_as_uint(s_v[hidx(x, y, z)]) ^ __float_as_uint(s_v[hidx(x2, y2, z2)])) >> 31) == 0) continue;
                f3 p, nrm;
                hermite(h, s_org, x, y, z, dir, p, nrm, &bad);
            }
        }
        // (c) is the cell whose corner 0 is this voxel a surface cell? (:187-199)
        bool valid = true; unsigned neg = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int i = hidx(x + c_corner[k][0], y + c_corner[k][1], z + c_corner[k][2]);
            valid = valid && s_a[i];
            neg |= (__float_as_uint(s_v[i]) >> 31) << k;
        }
        valid = valid && neg != 0 && neg != 255;
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, valid);
        if ((t & 31) == 0) s_bal[c >> 5] = bal;
    }
    if (bad && (!owned || owned[b])) flags[0] = 1;  // halo bricks of a sharded volume miss part of their own halo
    __syncthreads();
    if (t < 8) cell_valid[b * 8 + t] = (u64)s_bal[2 * t] | ((u64)s_bal[2 * t + 1] << 32);
    unsigned n_valid = 0;
#pragma unroll
    for (int w = 0; w < 16; ++w) {
        const unsigned m = s_bal[w];
        if (t < 32 && ((m >> t) & 1u)) s_list[n_valid + __popc(m & ((1u << t) - 1u))] = (unsigned short)(w * 32 + t);
        n_valid += __popc(m);
    }
    __syncthreads();
    for (unsigned i = t; i < n_valid; i += DC_TPB) {
        const unsigned c = s_list[i];
        const int x = (c >> 6) + 1, y = ((c >> 3) & 7) + 1, z = (c & 7) + 1;
        // the cell's Hermite samples, compacted in EDGE_OFFSETS order (thread-local arrays; the solve walks only the m samples
        // that exist -- typically 3 to 6 of the 12 edges)
        f3 pts[12], nng[12]; int m = 0;
        bool bad2 = false;  // edges owned by other bricks are checked by their owners
#pragma unroll 1
        for (int e = 0; e < 12; ++e) {
            f3 p, nn;
            if (hermite(h, s_org, x + c_edge[e][0], y + c_edge[e][1], z + c_edge[e][2], c_edge[e][3], p, nn, &bad2)) { pts[m] = p; nng[m] = nn; ++m; }
        }
        // find_feature_point (:429-458)
        const float fm = (float)m;
        f3 cc{0.f, 0.f, 0.f};
        for (int i = 0; i < m; ++i) cc = xadd(cc, pts[i]);
        cc = f3{xdiv(cc.x, fm), xdiv(cc.y, fm), xdiv(cc.z, fm)};
#pragma unroll 1
        for (int it = 0; it < 50; ++it) {
            f3 force{0.f, 0.f, 0.f};
#pragma unroll 1
            for (int i = 0; i < m; ++i) {
                const f3 nn = nng[i];
                const f3 nneg = xscale(nn, -1.0f);
                force = xadd(force, xscale(nneg, xdot(nn, xsub(cc, pts[i]))));
            }
            const f3 fd = xscale(force, s_damp[it]);
            cc = xadd(cc, f3{xdiv(fd.x, fm), xdiv(fd.y, fm), xdiv(fd.z, fm)});
            if (xnorm2(force) < 1e-6f) break;
        }
        float* o = cell_pts + (b * 512 + c) * 3;
        o[0] = cc.x; o[1] = cc.y; o[2] = cc.z;
    }
}

template <bool WRITE>
__global__ void __launch_bounds__(512) k_dc_quads(const u64* __restrict__ keys, const float* __restrict__ values, const u64* __restrict__ masks, size_t n,
                                                  const int* __restrict__ nbr, const unsigned char* __restrict__ owned, const float* __restrict__ cell_pts, const u64* __restrict__ cell_valid, float vs,
                                                  unsigned* counts, const u64* __restrict__ offsets, float* out) {
    __shared__ long long s_nb[8];   // bit0 = -x, bit1 = -y, bit2 = -z neighbour (cells), index 0 = this brick
    __shared__ long long s_pb[4];   // +x, +y, +z neighbour (values), index 0 = this brick
    const size_t b = blockIdx.x;
    const unsigned t = threadIdx.x;
    if (owned && !owned[b]) { if (!WRITE && t == 0) counts[b] = 0; return; }
    if (WRITE) { if (offsets[b + 1] == offsets[b]) return; }
    if (t < 8) s_nb[t] = nbr[b * 27 + (1 - (t & 1)) * 9 + (1 - ((t >> 1) & 1)) * 3 + (1 - ((t >> 2) & 1))];
    else if (t >= 32 && t < 36) { const unsigned d = t - 32; s_pb[d] = nbr[b * 27 + (d == 0 ? 13 : (d == 1 ? 22 : (d == 2 ? 16 : 14)))]; }
    __syncthreads();
    const int x = t >> 6, y = (t >> 3) & 7, z = t & 7;
    int ntri = 0;
    float local[WRITE ? 54 : 1];
    const bool act = (masks[b * 8 + (t >> 6)] >> (t & 63)) & 1;
    if (act) {
        const float v1 = values[b * 512 + t];
        for (int dir = 0; dir < 3; ++dir) {  // handle_edge (:99-135)
            const int x2 = x + (dir == 0), y2 = y + (dir == 1), z2 = z + (dir == 2);
            const long long vb = s_pb[(x2 > 7) ? 1 : ((y2 > 7) ? 2 : ((z2 > 7) ? 3 : 0))];
            if (vb < 0) continue;
            const unsigned off2 = ((x2 & 7) << 6) | ((y2 & 7) << 3) | (z2 & 7);
            if (!((masks[vb * 8 + (off2 >> 6)] >> (off2 & 63)) & 1)) continue;
            const float v2 = values[vb * 512 + off2];
            if (((__float_as_uint(v1) ^ __float_as_uint(v2)) >> 31) == 0) continue;
            f3 p[4]; bool ok = true;
            for (int k = 0; k < 4 && ok; ++k) {
                const int cx = x + c_cell[dir][k][0], cy = y + c_cell[dir][k][1], cz = z + c_cell[dir][k][2];
                const long long cb = s_nb[(cx < 0 ? 1 : 0) | (cy < 0 ? 2 : 0) | (cz < 0 ? 4 : 0)];
                if (cb < 0) { ok = false; break; }
                const unsigned co = ((cx & 7) << 6) | ((cy & 7) << 3) | (cz & 7);
                if (!((cell_valid[cb * 8 + (co >> 6)] >> (co & 63)) & 1)) { ok = false; break; }
                const float* q = cell_pts + ((size_t)cb * 512 + co) * 3;
                p[k] = f3{q[0], q[1], q[2]};
            }
            if (!ok) continue;
            f3 f[6] = {p[0], p[1], p[2], p[2], p[3], p[0]};
            if (__float_as_uint(v1) >> 31) { f3 tmp = f[1]; f[1] = f[2]; f[2] = tmp; tmp = f[4]; f[4] = f[5]; f[5] = tmp; }
            for (int k = 0; k < 2; ++k) {  // dual_contouring.rs:67-78
                const f3 a = xscale(f[3 * k], vs), bb = xscale(f[3 * k + 1], vs), cc = xscale(f[3 * k + 2], vs);
                if (xnorm2(xcross(xsub(bb, a), xsub(cc, a))) == 0.f) continue;
                if (WRITE) { float* o = local + 9 * ntri; o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = bb.x; o[4] = bb.y; o[5] = bb.z; o[6] = cc.x; o[7] = cc.y; o[8] = cc.z; }
                ++ntri;
            }
        }
    }
    typedef cub::BlockScan<int, 512> Scan;
    __shared__ typename Scan::TempStorage tmp;
    int excl, total;
    Scan(tmp).ExclusiveSum(ntri, excl, total);
    if (!WRITE) { if (t == 0) counts[b] = (unsigned)total; return; }
    float* dst = out + (offsets[b] + (u64)excl) * 9;
    for (int i = 0; i < ntri * 9; ++i) dst[i] = local[i];
}

__global__ void k_widen(const unsigned* in, u64* out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
    if (i == n) out[i] = 0;
}

}  // namespace

bs_status bs_ensure_out_verts(bs_context* ctx, size_t n_floats);

bs_status bs_dc_impl(const bs_volume* v, float voxel_size, const float** d_verts, size_t* n_verts) {
    bs_context* ctx = v->ctx;
    cudaStream_t st = ctx->stream;
    *d_verts = nullptr; *n_verts = 0;
    bs_marks_begin(ctx);
    if (v->n_tiles8 || v->n_tiles128) return bs_fail(ctx, BS_ERR_REFERENCE_PANICS, "dual contouring over active tiles: the reference hits todo!() (dual_contouring.rs:139,186,347)");
    const size_t n = v->n_bricks;
    if (n == 0) { bs_marks_end(ctx); return BS_OK; }
    float* d_cells = nullptr; u64 *d_valid = nullptr, *d_wide = nullptr, *d_off = nullptr; unsigned* d_counts = nullptr; int *d_flags = nullptr, *d_nbr = nullptr;
    BS_TRY(bs_alloc(ctx, &d_nbr, n * 27));
    bs_count_launch(), k_dc_neighbours<<<bs_blocks(n * 27, 256), 256, 0, st>>>(v->keys, n, d_nbr);
    BS_TRY(bs_alloc(ctx, &d_cells, n * 512 * 3)); BS_TRY(bs_alloc(ctx, &d_valid, n * 8)); BS_TRY(bs_alloc(ctx, &d_flags, 1));
    BS_TRY(bs_alloc(ctx, &d_counts, n)); BS_TRY(bs_alloc(ctx, &d_wide, n + 1)); BS_TRY(bs_alloc(ctx, &d_off, n + 1));
    BS_CUDA(ctx, cudaMemsetAsync(d_flags, 0, sizeof(int), st));
    bs_count_launch(), k_dc_cells<<<(unsigned)n, DC_TPB, 0, st>>>(v->keys, v->values, v->masks, n, d_nbr, v->owned, d_cells, d_valid, d_flags);
    bs_mark(ctx, "dc_cells_ms");
    bs_count_launch(), k_dc_quads<false><<<(unsigned)n, 512, 0, st>>>(v->keys, v->values, v->masks, n, d_nbr, v->owned, d_cells, d_valid, voxel_size, d_counts, nullptr, nullptr);
    bs_count_launch(), k_widen<<<bs_blocks(n + 1, 256), 256, 0, st>>>(d_counts, d_wide, n);
    void* d_tmp = nullptr; size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_wide, d_off, n + 1, st);
    BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp_bytes));
    cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_wide, d_off, n + 1, st);
    u64 n_tris = 0; int flag = 0;
    BS_TRY(bs_fetch(ctx, &n_tris, d_off + n, sizeof(n_tris)));
    BS_TRY(bs_fetch(ctx, &flag, d_flags, sizeof(int)));
    BS_TRY(bs_sync(ctx));
    bs_mark(ctx, "dc_count_ms");
    bs_status s = BS_OK;
    if (flag) s = bs_fail(ctx, BS_ERR_REFERENCE_PANICS, "dual contouring: a sign-change edge end point has no neighbour along some axis; the reference hits unreachable!() (dual_contouring.rs:340)");
    if (s == BS_OK) s = bs_ensure_out_verts(ctx, (size_t)n_tris * 9);
    if (s == BS_OK && n_tris) bs_count_launch(), k_dc_quads<true><<<(unsigned)n, 512, 0, st>>>(v->keys, v->values, v->masks, n, d_nbr, v->owned, d_cells, d_valid, voxel_size, nullptr, d_off, ctx->d_out_verts);
    bs_mark(ctx, "dc_emit_ms");
    bs_free(ctx, d_tmp); bs_free(ctx, d_cells); bs_free(ctx, d_valid); bs_free(ctx, d_counts); bs_free(ctx, d_wide); bs_free(ctx, d_off); bs_free(ctx, d_flags); bs_free(ctx, d_nbr);
    if (s != BS_OK) return s;
    BS_TRY(bs_sync(ctx));
    BS_CUDA(ctx, cudaGetLastError());
    bs_marks_end(ctx);
    bs_stat_add(ctx, "n_bricks", (double)n);
    bs_stat_add(ctx, "n_out_tris", (double)n_tris);
    *d_verts = ctx->d_out_verts; *n_verts = (size_t)n_tris * 3;
    return BS_OK;
}
