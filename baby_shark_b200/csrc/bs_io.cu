// The data formats either side of the path, on the device (sm_100a):
//   - binary STL <-> n x 9 f32 triangle soups      (replaces io/stl.rs:65-95 StlReader, :143-191 StlWriter)
//   - ActiveVoxelsMesher::mesh                      (replaces voxel/meshing/active_voxels.rs:12-177)
//   - merge_points                                  (replaces algo/merge_points.rs:12-41 + data_structures/vertex_index_map.rs)
// All three are byte / index work bound by HBM: coalesced 32-bit accesses through shared-memory staging for the
// 50-byte STL records, one pass + scan + one pass for the two compactions.
#include "bs_common.cuh"
#include <cub/cub.cuh>
#include <algorithm>

namespace {

typedef unsigned long long u64;
constexpr int STL_CH = 256;       // records per CTA
constexpr int STL_HDR = 84;       // 80-byte header + u32 count

// ---- STL decode: 50-byte records {normal, v1, v2, v3, attr} -> 9 floats ----------------------------------------------
// A CTA copies its 12 800 contiguous bytes into shared memory with coalesced 32-bit loads (84 + 12800 k is 4-byte
// aligned), then writes the 9216 output bytes with coalesced 32-bit stores; the 2-byte misalignment of every second
// record is absorbed by reading shared memory as halfwords.
__global__ void __launch_bounds__(STL_CH) k_stl_decode(const unsigned char* __restrict__ stl, size_t n_tris, float* __restrict__ tris) {
    __shared__ unsigned s_w[STL_CH * 50 / 4];
    const size_t r0 = (size_t)blockIdx.x * STL_CH;
    const unsigned nrec = (unsigned)min((size_t)STL_CH, n_tris - r0);
    const unsigned char* src = stl + STL_HDR + r0 * 50;
    const unsigned nbytes = nrec * 50, nfull = nbytes >> 2;
    if ((reinterpret_cast<uintptr_t>(src) & 3u) == 0) {
        const unsigned* s4 = reinterpret_cast<const unsigned*>(src);
        for (unsigned i = threadIdx.x; i < nfull; i += STL_CH) s_w[i] = s4[i];
        if (threadIdx.x == 0 && (nbytes & 3u)) s_w[nfull] = (unsigned)src[nfull * 4] | ((unsigned)src[nfull * 4 + 1] << 8);  // odd record count: 2 tail bytes
    } else {  // caller-provided buffer that is not word aligned
        unsigned char* sb = reinterpret_cast<unsigned char*>(s_w);
        for (unsigned i = threadIdx.x; i < nbytes; i += STL_CH) sb[i] = src[i];
    }
    __syncthreads();
    const unsigned short* sh = reinterpret_cast<const unsigned short*>(s_w);
    unsigned* dst = reinterpret_cast<unsigned*>(tris + r0 * 9);
    for (unsigned w = threadIdx.x; w < nrec * 9; w += STL_CH) {
        const unsigned rec = w / 9, f = w - rec * 9;
        const unsigned h = rec * 25 + 6 + f * 2;  // halfword index of byte rec*50 + 12 + 4 f
        dst[w] = (unsigned)sh[h] | ((unsigned)sh[h + 1] << 16);
    }
}

// ---- STL encode: soup -> records with the recomputed normal (Triangle3::normal, triangle3.rs:261-269) -----------------
__global__ void __launch_bounds__(STL_CH) k_stl_encode(const float* __restrict__ verts, size_t n_tris, unsigned char* __restrict__ stl) {
    __shared__ float s_f[STL_CH * 9];
    __shared__ unsigned s_w[STL_CH * 50 / 4 + 1];
    const size_t r0 = (size_t)blockIdx.x * STL_CH;
    const unsigned nrec = (unsigned)min((size_t)STL_CH, n_tris - r0);
    for (unsigned w = threadIdx.x; w < nrec * 9; w += STL_CH) s_f[w] = verts[r0 * 9 + w];
    if (blockIdx.x == 0) {  // header: 80 zero bytes + count (io/stl.rs:153-161)
        unsigned* h = reinterpret_cast<unsigned*>(stl);
        if (threadIdx.x < 20) h[threadIdx.x] = 0u;
        if (threadIdx.x == 20) h[20] = (unsigned)n_tris;
    }
    __syncthreads();
    if (threadIdx.x < nrec) {
        const float* p = s_f + threadIdx.x * 9;
        const f3 a{p[0], p[1], p[2]}, b{p[3], p[4], p[5]}, c{p[6], p[7], p[8]};
        const f3 cr = xcross(xsub(b, a), xsub(c, a));
        const float n2 = xnorm2(cr);
        f3 nrm{0.f, 0.f, 0.f};  // zeros for a degenerate face (io/stl.rs:171)
        if (!(n2 == 0.f)) { const float len = xsqrt(n2); nrm = f3{xdiv(cr.x, len), xdiv(cr.y, len), xdiv(cr.z, len)}; }  // cross.normalize() = cross / norm
        unsigned short* sh = reinterpret_cast<unsigned short*>(s_w) + threadIdx.x * 25;
        const float rec[12] = {nrm.x, nrm.y, nrm.z, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8]};
#pragma unroll
        for (int k = 0; k < 12; ++k) { const unsigned u = __float_as_uint(rec[k]); sh[2 * k] = (unsigned short)(u & 0xFFFFu); sh[2 * k + 1] = (unsigned short)(u >> 16); }
        sh[24] = 0;  // attribute byte count
    }
    if (threadIdx.x == 0) reinterpret_cast<unsigned short*>(s_w)[nrec * 25] = 0;  // pad the tail word of an odd chunk
    __syncthreads();
    unsigned* dst = reinterpret_cast<unsigned*>(stl + STL_HDR + r0 * 50);  // library-allocated: word aligned, padded
    for (unsigned i = threadIdx.x; i < (nrec * 50 + 3) / 4; i += STL_CH) dst[i] = s_w[i];
}

// ---- ActiveVoxelsMesher ------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long av_find_key(const u64* keys, size_t n, u64 k) {
    size_t lo = 0, hi = n;
    while (lo < hi) { size_t mid = (lo + hi) >> 1; if (keys[mid] < k) lo = mid + 1; else hi = mid; }
    return (lo < n && keys[lo] == k) ? (long long)lo : -1;
}
// One warp per brick. s_m[0] = the brick's mask words, s_m[1 + d] = those of the face neighbour d (+z -z -x +x +y -y,
// the reference's probe order top bottom left right front back), zeros when absent.
struct AvTiles { const u64* t8k; size_t nt8; const u64* t128k; size_t nt128; };
// active tile (8^3 keyed like a brick, 128^3 keyed by key >> 12) covering brick (bx, by, bz)?
__device__ __forceinline__ bool av_tile_covers(const AvTiles& Tl, int bx, int by, int bz) {
    if (bx < BS_BRICK_MIN || bx > BS_BRICK_MAX || by < BS_BRICK_MIN || by > BS_BRICK_MAX || bz < BS_BRICK_MIN || bz > BS_BRICK_MAX) return false;
    const u64 k = bs_brick_key(bx, by, bz);
    return (Tl.nt8 && av_find_key(Tl.t8k, Tl.nt8, k) >= 0) || (Tl.nt128 && av_find_key(Tl.t128k, Tl.nt128, k >> 12) >= 0);
}
template <bool WRITE>
__global__ void __launch_bounds__(128) k_active_voxels(const u64* __restrict__ keys, const u64* __restrict__ masks, size_t n, AvTiles Tl, const unsigned* __restrict__ pos, unsigned* counts, const u64* __restrict__ offsets, int* out) {
    __shared__ u64 s_m[4][7][8];
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const size_t b = (size_t)blockIdx.x * 4 + w;
    if (b >= n) return;
    int bx, by, bz; bs_key_brick(keys[b], bx, by, bz);
    const int N[6][3] = {{0, 0, 1}, {0, 0, -1}, {-1, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, -1, 0}};
    long long nb = -1;
    if (lane >= 1 && lane < 7) {
        const int d = lane - 1, x = bx + N[d][0], y = by + N[d][1], z = bz + N[d][2];
        if (x >= BS_BRICK_MIN && x <= BS_BRICK_MAX && y >= BS_BRICK_MIN && y <= BS_BRICK_MAX && z >= BS_BRICK_MIN && z <= BS_BRICK_MAX) nb = av_find_key(keys, n, bs_brick_key(x, y, z));
    }
    if (lane == 0) nb = (long long)b;
    if (lane >= 1 && lane < 7 && nb < 0 && (Tl.nt8 | Tl.nt128)) {  // no brick there: an active tile reads as all active (grid.at is Some)
        const int d = lane - 1;
        if (av_tile_covers(Tl, bx + N[d][0], by + N[d][1], bz + N[d][2])) nb = -2;
    }
    for (int k = 0; k < 7; ++k) {
        const long long src = __shfl_sync(0xFFFFFFFFu, nb, k);
        if (lane < 8) s_m[w][k][lane] = src >= 0 ? masks[(size_t)src * 8 + lane] : (src == -2 ? ~0ull : 0ull);
    }
    __syncwarp();
    auto active = [&](int x, int y, int z) -> bool {  // (x, y, z) in -1..8, at most one coordinate outside 0..7
        int k = 0;
        if (z > 7) k = 1; else if (z < 0) k = 2; else if (x < 0) k = 3; else if (x > 7) k = 4; else if (y > 7) k = 5; else if (y < 0) k = 6;
        return (s_m[w][k][x & 7] >> (((y & 7) << 3) | (z & 7))) & 1ull;
    };
    const int B[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};  // CUBE_OFFSETS (voxel/utils.rs:86-95)
    const int F[6][6] = {{4, 6, 7, 4, 5, 6}, {1, 0, 3, 1, 3, 2}, {0, 4, 3, 4, 7, 3}, {1, 6, 5, 1, 2, 6}, {2, 3, 6, 6, 3, 7}, {1, 5, 0, 5, 4, 0}};
    const size_t item = pos ? pos[b] : b;  // rank among bricks and tiles in visit order
    unsigned long long run = WRITE ? offsets[item] : 0ull;
    unsigned total = 0;
    for (int r = 0; r < 16; ++r) {  // voxels in leaf order x<<6 | y<<3 | z
        const unsigned o = r * 32 + lane;
        const int x = o >> 6, y = (o >> 3) & 7, z = o & 7;
        unsigned open = 0;
        if (active(x, y, z)) {
#pragma unroll
            for (int d = 0; d < 6; ++d) if (!active(x + N[d][0], y + N[d][1], z + N[d][2])) open |= 1u << d;
        }
        const unsigned cnt = __popc(open);
        unsigned inc = cnt;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { const unsigned t = __shfl_up_sync(0xFFFFFFFFu, inc, s); if ((int)lane >= s) inc += t; }
        if (WRITE && open) {
            int* dst = out + (run + (inc - cnt)) * 18;  // 6 vertices x 3 ints per open face
            const int vx = (bx << 3) + x, vy = (by << 3) + y, vz = (bz << 3) + z;
            for (int d = 0; d < 6; ++d) {
                if (!((open >> d) & 1)) continue;
                for (int k = 0; k < 6; ++k) { const int* c = B[F[d][k]]; dst[0] = vx + c[0]; dst[1] = vy + c[1]; dst[2] = vz + c[2]; dst += 3; }
            }
        }
        const unsigned rt = __shfl_sync(0xFFFFFFFFu, inc, 31);
        run += rt; total += rt;
    }
    if (!WRITE && lane == 0) counts[item] = total;  // open faces of the brick
}

// is voxel (x, y, z) active (in a brick with its bit set, or inside an active tile)?  TreeNode::at(...).is_some()
__device__ bool av_voxel_active(const u64* __restrict__ keys, const u64* __restrict__ masks, size_t n, const AvTiles& Tl, int x, int y, int z) {
    const int bx = x >> 3, by = y >> 3, bz = z >> 3;
    if (bx < BS_BRICK_MIN || bx > BS_BRICK_MAX || by < BS_BRICK_MIN || by > BS_BRICK_MAX || bz < BS_BRICK_MIN || bz > BS_BRICK_MAX) return false;
    const long long b = av_find_key(keys, n, bs_brick_key(bx, by, bz));
    if (b >= 0) return (masks[(size_t)b * 8 + (x & 7)] >> (((y & 7) << 3) | (z & 7))) & 1ull;
    return av_tile_covers(Tl, bx, by, bz);
}
// Active tiles (active_voxels.rs:128-151): only boundary voxels are tested, for every (i, j) in the order left, right,
// top, bottom, front, back -- edge and corner voxels several times, duplicates are emitted as the reference does.
// One CTA per tile; candidates (6 per (i, j)) are taken 256 at a time and block-scanned to keep that order.
template <bool WRITE>
__global__ void __launch_bounds__(256) k_active_voxels_tiles(const u64* __restrict__ keys, const u64* __restrict__ masks, size_t n, AvTiles Tl, int level /*0: 8^3, 1: 128^3*/,
                                                             const unsigned* __restrict__ pos, unsigned* counts, const u64* __restrict__ offsets, int* out) {
    typedef cub::BlockScan<unsigned, 256> Scan;
    __shared__ typename Scan::TempStorage tmp;
    const size_t t = blockIdx.x;
    const int size = level ? 128 : 8;
    int ox, oy, oz;
    {
        int bx, by, bz;
        if (level) bs_key_brick(Tl.t128k[t] << 12, bx, by, bz); else bs_key_brick(Tl.t8k[t], bx, by, bz);
        ox = bx << 3; oy = by << 3; oz = bz << 3;
    }
    const int N[6][3] = {{0, 0, 1}, {0, 0, -1}, {-1, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, -1, 0}};
    const int B[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
    const int F[6][6] = {{4, 6, 7, 4, 5, 6}, {1, 0, 3, 1, 3, 2}, {0, 4, 3, 4, 7, 3}, {1, 6, 5, 1, 2, 6}, {2, 3, 6, 6, 3, 7}, {1, 5, 0, 5, 4, 0}};
    const size_t item = pos[t];
    unsigned long long run = WRITE ? offsets[item] : 0ull;
    unsigned long long total = 0;
    const unsigned n_cand = 6u * (unsigned)size * (unsigned)size;
    for (unsigned base = 0; base < n_cand; base += 256) {
        const unsigned c = base + threadIdx.x;
        unsigned open = 0; int vx = 0, vy = 0, vz = 0;
        if (c < n_cand) {
            const int pair = (int)(c / 6), which = (int)(c % 6), i = pair / size, j = pair % size, s1 = size - 1;
            const int lx = which == 0 ? 0 : (which == 1 ? s1 : i), ly = which < 2 ? i : (which < 4 ? j : (which == 4 ? s1 : 0)), lz = which < 2 ? j : (which == 2 ? s1 : (which == 3 ? 0 : j));
            vx = ox + lx; vy = oy + ly; vz = oz + lz;
            for (int d = 0; d < 6; ++d) {
                const int qx = lx + N[d][0], qy = ly + N[d][1], qz = lz + N[d][2];
                const bool inside = qx >= 0 && qx < size && qy >= 0 && qy < size && qz >= 0 && qz < size;
                if (!inside && !av_voxel_active(keys, masks, n, Tl, ox + qx, oy + qy, oz + qz)) open |= 1u << d;
            }
        }
        const unsigned cnt = __popc(open);
        unsigned excl, sum;
        Scan(tmp).ExclusiveSum(cnt, excl, sum);
        if (WRITE && open) {
            int* dst = out + (run + excl) * 18;
            for (int d = 0; d < 6; ++d) {
                if (!((open >> d) & 1)) continue;
                for (int k = 0; k < 6; ++k) { const int* cc = B[F[d][k]]; dst[0] = vx + cc[0]; dst[1] = vy + cc[1]; dst[2] = vz + cc[2]; dst += 3; }
            }
        }
        run += sum; total += sum;
        __syncthreads();
    }
    if (!WRITE && threadIdx.x == 0) counts[item] = (unsigned)total;
}
// rank of every brick / tile in the merged key order (a 128^3 tile sorts by its 16^3-node key), as in bs_mc.cu
__global__ void k_av_merge_pos(const u64* __restrict__ keys, size_t n, AvTiles Tl, unsigned* pos_brick, unsigned* pos_t8, unsigned* pos_t128) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    auto lb = [](const u64* k, size_t m, u64 key, int shift) { size_t lo = 0, hi = m; while (lo < hi) { const size_t mid = (lo + hi) >> 1; if ((k[mid] << shift) < key) lo = mid + 1; else hi = mid; } return (unsigned)lo; };
    if (i < n) pos_brick[i] = (unsigned)i + lb(Tl.t8k, Tl.nt8, keys[i], 0) + lb(Tl.t128k, Tl.nt128, keys[i], 12);
    if (i < Tl.nt8) pos_t8[i] = (unsigned)i + lb(keys, n, Tl.t8k[i], 0) + lb(Tl.t128k, Tl.nt128, Tl.t8k[i], 12);
    if (i < Tl.nt128) pos_t128[i] = (unsigned)i + lb(keys, n, Tl.t128k[i] << 12, 0) + lb(Tl.t8k, Tl.nt8, Tl.t128k[i] << 12, 0);
}
__global__ void k_widen_u32(const unsigned* in, u64* out, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
    if (i == n) out[i] = 0;
}

// ---- merge_points ------------------------------------------------------------------------------------------------------
constexpr unsigned MP_EMPTY = 0xFFFFFFFFu;
__device__ __forceinline__ unsigned mp_hash(float x, float y, float z) {
    // +0.0f folds -0 onto +0, which compare equal in the reference (PartialEq on f32)
    const unsigned a = __float_as_uint(x + 0.0f), b = __float_as_uint(y + 0.0f), c = __float_as_uint(z + 0.0f);
    unsigned h = a * 73856093u ^ b * 19349663u ^ c * 83492791u;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
    return h;
}
// Open addressing, one word per slot = the smallest index seen so far of the slot's point class; the key is read back
// from the point array. NaN never compares equal, so a NaN point ends up alone in its own slot, as in the reference.
__global__ void k_mp_insert(const float* __restrict__ pts, size_t n, unsigned* table, unsigned mask, unsigned* slot_of) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    unsigned s = mp_hash(x, y, z) & mask;
    for (;;) {
        unsigned cur = table[s];
        if (cur == MP_EMPTY) { cur = atomicCAS(table + s, MP_EMPTY, (unsigned)i); if (cur == MP_EMPTY) break; }
        if (pts[3 * (size_t)cur] == x && pts[3 * (size_t)cur + 1] == y && pts[3 * (size_t)cur + 2] == z) { atomicMin(table + s, (unsigned)i); break; }
        s = (s + 1) & mask;
    }
    slot_of[i] = s;
}
__global__ void k_mp_first(const unsigned* __restrict__ table, const unsigned* __restrict__ slot_of, size_t n, unsigned* first, unsigned* is_first) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned f = table[slot_of[i]];
    first[i] = f; is_first[i] = f == (unsigned)i;
}
__global__ void k_mp_emit(const float* __restrict__ pts, const unsigned* __restrict__ first, const unsigned* __restrict__ rank_excl, size_t n, float* unique, unsigned* indices) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned f = first[i], r = rank_excl[f];
    indices[i] = r;
    if (f == (unsigned)i) { unique[3 * (size_t)r] = pts[3 * i]; unique[3 * (size_t)r + 1] = pts[3 * i + 1]; unique[3 * (size_t)r + 2] = pts[3 * i + 2]; }
}

}  // namespace

// ---- entry points (bshark.h) -----------------------------------------------------------------------------------------------
bs_status bs_stl_decode_impl(bs_context* ctx, const unsigned char* d_stl, size_t n_bytes, float** d_tris, size_t* n_tris) {
    cudaStream_t st = ctx->stream;
    *d_tris = nullptr; *n_tris = 0;
    if (n_bytes < (size_t)STL_HDR) return bs_fail(ctx, BS_ERR_INVALID, "STL: %zu bytes is shorter than the 84-byte header (the reference fails with ReadError)", n_bytes);
    unsigned count = 0;
    BS_CUDA(ctx, cudaMemcpyAsync(&count, d_stl + 80, 4, cudaMemcpyDeviceToHost, st));
    BS_CUDA(ctx, cudaStreamSynchronize(st));
    if (n_bytes < (size_t)STL_HDR + (size_t)count * 50) return bs_fail(ctx, BS_ERR_INVALID, "STL: header announces %u triangles but only %zu bytes follow (the reference fails with ReadError)", count, n_bytes - STL_HDR);
    float* t = nullptr;
    BS_TRY(bs_alloc(ctx, &t, (size_t)count * 9));
    bs_marks_begin(ctx);
    if (count) bs_count_launch(), k_stl_decode<<<bs_blocks(count, STL_CH), STL_CH, 0, st>>>(d_stl, count, t);
    bs_mark(ctx, "stl_decode_ms");
    BS_CUDA(ctx, cudaGetLastError());
    bs_marks_end(ctx);
    bs_stat_add(ctx, "n_tris", (double)count);
    *d_tris = t; *n_tris = count;
    return BS_OK;
}

bs_status bs_stl_encode_impl(bs_context* ctx, const float* d_verts, size_t n_verts, unsigned char** d_stl, size_t* n_bytes) {
    cudaStream_t st = ctx->stream;
    *d_stl = nullptr; *n_bytes = 0;
    if (n_verts % 3) return bs_fail(ctx, BS_ERR_INVALID, "STL: a soup has 3 vertices per triangle");
    const size_t n_tris = n_verts / 3;
    if (n_tris > 0xFFFFFFFFull) return bs_fail(ctx, BS_ERR_RANGE, "Mesh is too big for STL");  // io/stl.rs:156-158
    unsigned char* o = nullptr;
    BS_TRY(bs_alloc(ctx, &o, STL_HDR + n_tris * 50 + 4));
    bs_marks_begin(ctx);
    if (n_tris) bs_count_launch(), k_stl_encode<<<bs_blocks(n_tris, STL_CH), STL_CH, 0, st>>>(d_verts, n_tris, o);
    else { BS_CUDA(ctx, cudaMemsetAsync(o, 0, STL_HDR, st)); }
    bs_mark(ctx, "stl_encode_ms");
    BS_CUDA(ctx, cudaGetLastError());
    bs_marks_end(ctx);
    bs_stat_add(ctx, "n_tris", (double)n_tris);
    *d_stl = o; *n_bytes = STL_HDR + n_tris * 50;
    return BS_OK;
}

bs_status bs_active_voxels_impl(const bs_volume* v, int** d_verts, size_t* n_verts) {
    bs_context* ctx = v->ctx;
    cudaStream_t st = ctx->stream;
    *d_verts = nullptr; *n_verts = 0;
    const size_t n = v->n_bricks, nt8 = v->n_tiles8, nt128 = v->n_tiles128, n_items = n + nt8 + nt128;
    if (n_items == 0) return BS_OK;
    bs_marks_begin(ctx);
    const AvTiles Tl{v->tile8_keys, nt8, v->tile128_keys, nt128};
    unsigned *d_cnt = nullptr, *d_pos = nullptr; u64 *d_wide = nullptr, *d_off = nullptr; void* d_tmp = nullptr; size_t tmp = 0;
    BS_TRY(bs_alloc(ctx, &d_cnt, n_items)); BS_TRY(bs_alloc(ctx, &d_wide, n_items + 1)); BS_TRY(bs_alloc(ctx, &d_off, n_items + 1));
    if (nt8 || nt128) {
        BS_TRY(bs_alloc(ctx, &d_pos, n_items));
        bs_count_launch(), k_av_merge_pos<<<bs_blocks(std::max(n, std::max(nt8, nt128)), 256), 256, 0, st>>>(v->keys, n, Tl, d_pos, d_pos + n, d_pos + n + nt8);
    }
    if (n) bs_count_launch(), k_active_voxels<false><<<bs_blocks(n, 4), 128, 0, st>>>(v->keys, v->masks, n, Tl, d_pos, d_cnt, nullptr, nullptr);
    if (nt8) bs_count_launch(), k_active_voxels_tiles<false><<<(unsigned)nt8, 256, 0, st>>>(v->keys, v->masks, n, Tl, 0, d_pos + n, d_cnt, nullptr, nullptr);
    if (nt128) bs_count_launch(), k_active_voxels_tiles<false><<<(unsigned)nt128, 256, 0, st>>>(v->keys, v->masks, n, Tl, 1, d_pos + n + nt8, d_cnt, nullptr, nullptr);
    bs_count_launch(), k_widen_u32<<<bs_blocks(n_items + 1, 256), 256, 0, st>>>(d_cnt, d_wide, n_items);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_wide, d_off, n_items + 1, st);
    BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp));
    cub::DeviceScan::ExclusiveSum(d_tmp, tmp, d_wide, d_off, n_items + 1, st);
    u64 faces = 0;
    BS_CUDA(ctx, cudaMemcpyAsync(&faces, d_off + n_items, sizeof(faces), cudaMemcpyDeviceToHost, st));
    BS_CUDA(ctx, cudaStreamSynchronize(st));
    int* o = nullptr;
    bs_status s = bs_alloc(ctx, &o, (size_t)faces * 18);
    if (s == BS_OK && faces) {
        if (n) bs_count_launch(), k_active_voxels<true><<<bs_blocks(n, 4), 128, 0, st>>>(v->keys, v->masks, n, Tl, d_pos, nullptr, d_off, o);
        if (nt8) bs_count_launch(), k_active_voxels_tiles<true><<<(unsigned)nt8, 256, 0, st>>>(v->keys, v->masks, n, Tl, 0, d_pos + n, nullptr, d_off, o);
        if (nt128) bs_count_launch(), k_active_voxels_tiles<true><<<(unsigned)nt128, 256, 0, st>>>(v->keys, v->masks, n, Tl, 1, d_pos + n + nt8, nullptr, d_off, o);
    }
    bs_mark(ctx, "active_voxels_ms");
    bs_free(ctx, d_tmp); bs_free(ctx, d_cnt); bs_free(ctx, d_wide); bs_free(ctx, d_off); bs_free(ctx, d_pos);
    if (s != BS_OK) return s;
    BS_CUDA(ctx, cudaGetLastError());
    bs_marks_end(ctx);
    *d_verts = o; *n_verts = (size_t)faces * 6;
    return BS_OK;
}

bs_status bs_merge_points_impl(bs_context* ctx, const float* d_pts, size_t n, float** d_unique, size_t* n_unique, unsigned** d_indices) {
    cudaStream_t st = ctx->stream;
    *d_unique = nullptr; *n_unique = 0; *d_indices = nullptr;
    if (n >= 0x7FFFFFFFull) return bs_fail(ctx, BS_ERR_RANGE, "merge_points: more than 2^31 - 1 points");
    unsigned* idx = nullptr;
    BS_TRY(bs_alloc(ctx, &idx, n));
    if (n == 0) { *d_indices = idx; return bs_alloc(ctx, d_unique, 1); }
    bs_marks_begin(ctx);
    size_t cap = 1024; while (cap < 2 * n) cap <<= 1;
    unsigned *table = nullptr, *slot_of = nullptr, *first = nullptr, *is_first = nullptr, *rank = nullptr; void* d_tmp = nullptr; size_t tmp = 0;
    BS_TRY(bs_alloc(ctx, &table, cap)); BS_TRY(bs_alloc(ctx, &slot_of, n)); BS_TRY(bs_alloc(ctx, &first, n)); BS_TRY(bs_alloc(ctx, &is_first, n + 1)); BS_TRY(bs_alloc(ctx, &rank, n + 1));
    BS_CUDA(ctx, cudaMemsetAsync(table, 0xFF, cap * sizeof(unsigned), st));
    BS_CUDA(ctx, cudaMemsetAsync(is_first + n, 0, sizeof(unsigned), st));
    bs_count_launch(), k_mp_insert<<<bs_blocks(n, 256), 256, 0, st>>>(d_pts, n, table, (unsigned)(cap - 1), slot_of);
    bs_count_launch(), k_mp_first<<<bs_blocks(n, 256), 256, 0, st>>>(table, slot_of, n, first, is_first);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, is_first, rank, n + 1, st);
    BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp));
    cub::DeviceScan::ExclusiveSum(d_tmp, tmp, is_first, rank, n + 1, st);
    unsigned nu = 0;
    BS_CUDA(ctx, cudaMemcpyAsync(&nu, rank + n, sizeof(nu), cudaMemcpyDeviceToHost, st));
    BS_CUDA(ctx, cudaStreamSynchronize(st));
    float* uq = nullptr;
    bs_status s = bs_alloc(ctx, &uq, (size_t)nu * 3);
    if (s == BS_OK) bs_count_launch(), k_mp_emit<<<bs_blocks(n, 256), 256, 0, st>>>(d_pts, first, rank, n, uq, idx);
    bs_mark(ctx, "merge_points_ms");
    bs_free(ctx, d_tmp); bs_free(ctx, table); bs_free(ctx, slot_of); bs_free(ctx, first); bs_free(ctx, is_first); bs_free(ctx, rank);
    if (s != BS_OK) { bs_free(ctx, idx); return s; }
    BS_CUDA(ctx, cudaGetLastError());
    bs_marks_end(ctx);
    bs_stat_add(ctx, "n_points", (double)n);
    bs_stat_add(ctx, "n_unique", (double)nu);
    *d_unique = uq; *n_unique = nu; *d_indices = idx;
    return BS_OK;
}
