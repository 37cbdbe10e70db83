// Inside/outside signs by fast winding numbers over a device-built BVH (sm_100a).
// Replaces WindingNumbers::{from_mesh, approximate} (src/spatial_partitioning/aabb_tree.rs:636-691,711-816)
// and MeshToVolume::compute_sings / ComputeSignsVisitor (src/voxel/mesh_to_volume.rs:198-281):
//     value = copysign(|d|, wn(p) < 0.2 ? +1 : -1),   p = idx as f32 * voxel_size.
//
// The reference walks a serial top-down SAH tree with 3 triangles per leaf. Here the hierarchy is an LBVH in
// its implicit form: triangles are radix-sorted by the 63-bit Morton code of their centroid, every LEAF
// consecutive triangles form a leaf and every 8 consecutive nodes of a level form a node of the next, so a
// node's children are contiguous (one 512 B coalesced group) and no child pointers or build-time atomics are
// needed. Per node the same quantities as the reference (aabb_tree.rs:723-801): area-weighted normal (order 1),
// sum(area * c * n^T) - p~ (sum area*n)^T (order 2), dipole centre p~ = sum(area*c)/sum(area); the radius is the
// distance from p~ to the farthest corner of the node's box (the reference takes the farther of the two
// extreme corners only, :734-736; this one is a true bound).
//
// Traversal is warp-cooperative: a warp owns 32 active voxels of ONE brick and walks ONE shared stack. A node
// is accepted as a far-field dipole only when it is far (|p - p~| > 2 radius, :666) for all 32 voxels, otherwise
// the whole warp descends; lanes that could have stopped earlier just get a more accurate sum. There is no
// divergence and every node read is a warp-uniform (broadcast) 16 B load.
// Only the 0.2 threshold matters downstream, so FMA contraction is allowed here (unlike the distance stage).
#include "bs_common.cuh"
#include <cub/cub.cuh>

namespace {

constexpr int LEAF = 4;           // triangles per leaf
constexpr int FAN = 8;            // children per internal node
constexpr int MAX_LEVELS = 12;    // 4 * 8^11 triangles
constexpr int STACK = 128;        // per-warp stack entries (<= 7 * levels + 8 live entries)
constexpr int WARPS_PER_BLOCK = 4;
constexpr float BETA = 2.0f;      // accuracy_scale (mesh_to_volume.rs:264)
constexpr float INV_4PI = 0.07957747154594767f;

struct Raw {  // additive moments of a node (aabb_tree.rs:694-700) + box
    float area, awc[3], awn[3], o1sum[9], bbmin[3], bbmax[3], pad[2];
};
static_assert(sizeof(Raw) == 96, "Raw is 24 floats");

struct Tree {
    const float4* nodes;    // 4 x float4 per node: {c.xyz, beta^2 r^2} {o1.xyz, trM} {m00 m11 m22 m01+m10} {m02+m20 m12+m21 - -}
    const float4* tris;     // 3 x float4 per sorted triangle (9 floats + pad), LEAF per leaf, padded with degenerate triangles
    unsigned level_off[MAX_LEVELS];
    unsigned level_cnt[MAX_LEVELS];
    int levels;             // root is level levels-1, index 0
};

__device__ __forceinline__ int f2ord(float f) { int b = __float_as_int(f); return b >= 0 ? b : b ^ 0x7FFFFFFF; }
__device__ __forceinline__ float ord2f(int o) { return __int_as_float(o >= 0 ? o : o ^ 0x7FFFFFFF); }

__global__ void k_centroid_bounds(const float* __restrict__ tris, size_t n, int* bounds /*min xyz, max xyz as ordered ints*/) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    float c[3] = {0, 0, 0}; bool ok = false;
    if (t < n) {
        const float* p = tris + 9 * t;
        for (int d = 0; d < 3; ++d) c[d] = (p[d] + p[3 + d] + p[6 + d]) * (1.0f / 3.0f);
        ok = (c[0] == c[0]) && (c[1] == c[1]) && (c[2] == c[2]) && fabsf(c[0]) < 1e30f && fabsf(c[1]) < 1e30f && fabsf(c[2]) < 1e30f;
    }
    for (int d = 0; d < 3; ++d) {
        int lo = ok ? f2ord(c[d]) : 0x7FFFFFFF, hi = ok ? f2ord(c[d]) : (int)0x80000000;
        for (int o = 16; o; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, o)); hi = max(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, o)); }
        if ((threadIdx.x & 31) == 0) { atomicMin(bounds + d, lo); atomicMax(bounds + 3 + d, hi); }
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
        unsigned long long q = (unsigned long long)(u * 2097151.0f);
        code |= spread21(q) << (2 - d);
    }
    codes[t] = code; ids[t] = (unsigned)t;
}

// level 0: gather LEAF sorted triangles, store them for traversal, accumulate the leaf's moments (aabb_tree.rs:749-777)
__global__ void k_leaves(const float* __restrict__ tris, const unsigned* __restrict__ ids, size_t n, float4* sorted, Raw* raw, unsigned n_leaves) {
    unsigned l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_leaves) return;
    Raw r;
    r.area = 0.f;
    for (int d = 0; d < 3; ++d) { r.awc[d] = 0.f; r.awn[d] = 0.f; r.bbmin[d] = 3.0e38f; r.bbmax[d] = -3.0e38f; }
    for (int i = 0; i < 9; ++i) r.o1sum[i] = 0.f;
    r.pad[0] = r.pad[1] = 0.f;
    for (int k = 0; k < LEAF; ++k) {
        size_t s = (size_t)l * LEAF + k;
        float v[9];
        if (s < n) { const float* p = tris + 9 * (size_t)ids[s]; for (int i = 0; i < 9; ++i) v[i] = p[i]; }
        else { for (int i = 0; i < 9; ++i) v[i] = 0.f; }  // padding: degenerate triangle, contributes nothing
        sorted[3 * s + 0] = make_float4(v[0], v[1], v[2], v[3]);
        sorted[3 * s + 1] = make_float4(v[4], v[5], v[6], v[7]);
        sorted[3 * s + 2] = make_float4(v[8], 0.f, 0.f, 0.f);
        if (s >= n) continue;
        for (int d = 0; d < 3; ++d) {
            r.bbmin[d] = fminf(r.bbmin[d], fminf(v[d], fminf(v[3 + d], v[6 + d])));
            r.bbmax[d] = fmaxf(r.bbmax[d], fmaxf(v[d], fmaxf(v[3 + d], v[6 + d])));
        }
        float e1[3] = {v[3] - v[0], v[4] - v[1], v[5] - v[2]}, e2[3] = {v[6] - v[0], v[7] - v[1], v[8] - v[2]};
        float cr[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        float n2 = cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2];
        if (!(n2 > 0.f)) continue;  // degenerate triangles are skipped (:757-759)
        float len = sqrtf(n2), area = 0.5f * len;
        float nn[3] = {cr[0] / len, cr[1] / len, cr[2] / len};
        float c[3] = {(v[0] + v[3] + v[6]) / 3.0f, (v[1] + v[4] + v[7]) / 3.0f, (v[2] + v[5] + v[8]) / 3.0f};
        r.area += area;
        for (int d = 0; d < 3; ++d) { r.awn[d] += area * nn[d]; r.awc[d] += area * c[d]; }
        for (int col = 0; col < 3; ++col) for (int row = 0; row < 3; ++row) r.o1sum[col * 3 + row] += area * c[row] * nn[col];
    }
    raw[l] = r;
}

__global__ void k_level_up(const Raw* __restrict__ child, unsigned n_child, Raw* parent, unsigned n_parent) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_parent) return;
    Raw r = child[(size_t)i * FAN];
    for (int k = 1; k < FAN; ++k) {
        size_t c = (size_t)i * FAN + k;
        if (c >= n_child) break;
        const Raw& q = child[c];
        r.area += q.area;
        for (int d = 0; d < 3; ++d) { r.awc[d] += q.awc[d]; r.awn[d] += q.awn[d]; r.bbmin[d] = fminf(r.bbmin[d], q.bbmin[d]); r.bbmax[d] = fmaxf(r.bbmax[d], q.bbmax[d]); }
        for (int j = 0; j < 9; ++j) r.o1sum[j] += q.o1sum[j];
    }
    parent[i] = r;
}

__global__ void k_finalize_nodes(const Raw* __restrict__ raw, unsigned n, float4* nodes) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Raw r = raw[i];
    float c[3], rad2 = 0.f;
    for (int d = 0; d < 3; ++d) {
        c[d] = r.awc[d] / r.area;  // NaN for an all-degenerate node: never "far", always descended (as in the reference)
        float e = fmaxf(fabsf(r.bbmin[d] - c[d]), fabsf(r.bbmax[d] - c[d]));
        rad2 += e * e;
    }
    float m[9];
    for (int col = 0; col < 3; ++col) for (int row = 0; row < 3; ++row) m[col * 3 + row] = r.o1sum[col * 3 + row] - c[row] * r.awn[col];
    // m[col*3+row]; symmetric combinations for r^T M r
    nodes[4 * (size_t)i + 0] = make_float4(c[0], c[1], c[2], BETA * BETA * rad2);
    nodes[4 * (size_t)i + 1] = make_float4(r.awn[0], r.awn[1], r.awn[2], m[0] + m[4] + m[8]);
    nodes[4 * (size_t)i + 2] = make_float4(m[0], m[4], m[8], m[1] + m[3]);
    nodes[4 * (size_t)i + 3] = make_float4(m[2] + m[6], m[5] + m[7], 0.f, 0.f);
}

// far-field dipole (aabb_tree.rs:667-669, hessians :803-816): o1 . r/(4 pi |r|^3) + M : (I/(4 pi |r|^3) - 3 r r^T/(4 pi |r|^5))
__device__ __forceinline__ float far_field(const float4* __restrict__ nd, float rx, float ry, float rz, float r2) {
    const float4 a = __ldg(nd + 1), b = __ldg(nd + 2), c = __ldg(nd + 3);
    const float inv_r = rsqrtf(r2);
    const float k = INV_4PI * inv_r * inv_r * inv_r;
    const float rMr = b.x * rx * rx + b.y * ry * ry + b.z * rz * rz + b.w * rx * ry + c.x * rx * rz + c.y * ry * rz;
    return k * ((a.x * rx + a.y * ry + a.z * rz) + a.w - 3.0f * rMr * (inv_r * inv_r));
}

// solid_angle / (4 pi) (aabb_tree.rs:582-628), Van Oosterom-Strackee form
__device__ __forceinline__ float tri_winding(const float4* __restrict__ t, float qx, float qy, float qz) {
    const float4 t0 = __ldg(t), t1 = __ldg(t + 1), t2 = __ldg(t + 2);
    const float ax = t0.x - qx, ay = t0.y - qy, az = t0.z - qz;
    const float bx = t0.w - qx, by = t1.x - qy, bz = t1.y - qz;
    const float cx = t1.z - qx, cy = t1.w - qy, cz = t2.x - qz;
    const float la = sqrtf(ax * ax + ay * ay + az * az), lb = sqrtf(bx * bx + by * by + bz * bz), lc = sqrtf(cx * cx + cy * cy + cz * cz);
    if (la == 0.f || lb == 0.f || lc == 0.f) return 0.f;
    const float det = ax * (by * cz - bz * cy) + ay * (bz * cx - bx * cz) + az * (bx * cy - by * cx);
    if (det == 0.f) return 0.f;
    const float den = la * lb * lc + (ax * bx + ay * by + az * bz) * lc + (ax * cx + ay * cy + az * cz) * lb + (bx * cx + by * cy + bz * cz) * la;
    return atan2f(det, den) * (2.0f * INV_4PI);
}

// Warp-cooperative winding number of 32 query points (one per lane). `stack` is this warp's shared-memory stack.
__device__ float warp_winding(const Tree& T, float qx, float qy, float qz, unsigned* stack) {
    const unsigned lane = threadIdx.x & 31;
    float wn = 0.f;
    int sp = 0;
    // visit(level, idx): far for all lanes -> accumulate; else leaf -> exact; else push
    auto visit = [&](int level, unsigned idx) {
        const float4* nd = T.nodes + 4 * (size_t)(T.level_off[level] + idx);
        const float4 h = __ldg(nd);
        const float rx = h.x - qx, ry = h.y - qy, rz = h.z - qz;
        const float r2 = rx * rx + ry * ry + rz * rz;
        if (__all_sync(0xFFFFFFFFu, r2 > h.w)) { wn += far_field(nd, rx, ry, rz, r2); return; }
        if (level == 0) {
            const float4* t = T.tris + 3 * (size_t)idx * LEAF;
#pragma unroll
            for (int k = 0; k < LEAF; ++k) wn += tri_winding(t + 3 * k, qx, qy, qz);
            return;
        }
        if (lane == 0) stack[sp] = ((unsigned)level << 28) | idx;
        ++sp;
    };
    visit(T.levels - 1, 0);
    while (sp > 0) {
        --sp;
        __syncwarp();
        const unsigned e = stack[sp];
        __syncwarp();
        const int level = (int)(e >> 28) - 1;
        const unsigned first = (e & 0x0FFFFFFFu) * FAN;
        const unsigned cnt = min((unsigned)FAN, T.level_cnt[level] - first);
        for (unsigned k = 0; k < cnt; ++k) visit(level, first + k);
    }
    return wn;
}

__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK) k_sign(Tree T, float* values, unsigned long long* masks, size_t n_bricks,
                                                               const unsigned long long* __restrict__ keys, float vs) {
    __shared__ unsigned s_stack[WARPS_PER_BLOCK][STACK];
    __shared__ unsigned short s_list[WARPS_PER_BLOCK][512];
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const size_t warp_global = (size_t)blockIdx.x * WARPS_PER_BLOCK + w, n_warps = (size_t)gridDim.x * WARPS_PER_BLOCK;
    for (size_t b = warp_global; b < n_bricks; b += n_warps) {
        float* bv = values + b * 512;
        // active list + mask words
        unsigned cnt = 0;
        unsigned my_mask_lo = 0, my_mask_hi = 0;
        for (int r = 0; r < 16; ++r) {
            const unsigned off = r * 32 + lane;
            const bool act = __float_as_uint(bv[off]) != BS_UDF_SENTINEL_BITS;
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, act);
            if (act) s_list[w][cnt + __popc(bal & ((1u << lane) - 1))] = (unsigned short)off;
            cnt += __popc(bal);
            if ((int)lane == (r >> 1)) { if (r & 1) my_mask_hi = bal; else my_mask_lo = bal; }
        }
        if (lane < 8) masks[b * 8 + lane] = (unsigned long long)my_mask_lo | ((unsigned long long)my_mask_hi << 32);
        __syncwarp();
        if (cnt == 0) continue;
        int bx, by, bz; bs_key_brick(keys[b], bx, by, bz);
        for (unsigned base = 0; base < cnt; base += 32) {
            const unsigned i = base + lane;
            const bool valid = i < cnt;
            const unsigned off = s_list[w][valid ? i : cnt - 1];  // idle lanes shadow a real voxel: decisions stay brick-local
            const int x = (bx << 3) + (int)(off >> 6), y = (by << 3) + (int)((off >> 3) & 7), z = (bz << 3) + (int)(off & 7);
            const float wn = warp_winding(T, __fmul_rn((float)x, vs), __fmul_rn((float)y, vs), __fmul_rn((float)z, vs), s_stack[w]);
            if (valid) {
                const float d = bv[off];
                bv[off] = (wn < 0.2f) ? copysignf(d, 1.0f) : copysignf(d, -1.0f);  // mesh_to_volume.rs:266-271
            }
        }
        __syncwarp();
    }
}

}  // namespace

bs_status bs_sign_impl(bs_context* ctx, const float* d_tris, size_t n_tris, bs_volume* vol) {
    cudaStream_t st = ctx->stream;
    if (n_tris >= (1ull << 28) * LEAF) return bs_fail(ctx, BS_ERR_RANGE, "too many triangles");
    // Morton order
    int* d_bounds = nullptr; unsigned long long *d_codes = nullptr, *d_codes2 = nullptr; unsigned *d_ids = nullptr, *d_ids2 = nullptr;
    BS_TRY(bs_alloc(ctx, &d_bounds, 6));
    const int init[6] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF, (int)0x80000000, (int)0x80000000, (int)0x80000000};
    BS_CUDA(ctx, cudaMemcpyAsync(d_bounds, init, sizeof(init), cudaMemcpyHostToDevice, st));
    BS_TRY(bs_alloc(ctx, &d_codes, n_tris)); BS_TRY(bs_alloc(ctx, &d_codes2, n_tris));
    BS_TRY(bs_alloc(ctx, &d_ids, n_tris)); BS_TRY(bs_alloc(ctx, &d_ids2, n_tris));
    k_centroid_bounds<<<bs_blocks(n_tris, 256), 256, 0, st>>>(d_tris, n_tris, d_bounds);
    k_morton<<<bs_blocks(n_tris, 256), 256, 0, st>>>(d_tris, n_tris, d_bounds, d_codes, d_ids);
    void* d_tmp = nullptr; size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_codes, d_codes2, d_ids, d_ids2, n_tris, 0, 63, st);
    BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp_bytes));
    cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_codes, d_codes2, d_ids, d_ids2, n_tris, 0, 63, st);
    bs_free(ctx, d_tmp); bs_free(ctx, d_codes); bs_free(ctx, d_codes2); bs_free(ctx, d_ids); bs_free(ctx, d_bounds);
    // implicit hierarchy
    Tree T;
    unsigned cnt = (unsigned)((n_tris + LEAF - 1) / LEAF), total_nodes = 0;
    T.levels = 0;
    for (;;) {
        if (T.levels >= MAX_LEVELS) return bs_fail(ctx, BS_ERR_RANGE, "BVH too deep");
        T.level_off[T.levels] = total_nodes; T.level_cnt[T.levels] = cnt; total_nodes += cnt; ++T.levels;
        if (cnt == 1) break;
        cnt = (cnt + FAN - 1) / FAN;
    }
    for (int l = T.levels; l < MAX_LEVELS; ++l) { T.level_off[l] = 0; T.level_cnt[l] = 0; }
    float4 *d_sorted = nullptr, *d_nodes = nullptr; Raw* d_raw = nullptr;
    BS_TRY(bs_alloc(ctx, &d_sorted, (size_t)T.level_cnt[0] * LEAF * 3));
    BS_TRY(bs_alloc(ctx, &d_nodes, (size_t)total_nodes * 4));
    BS_TRY(bs_alloc(ctx, &d_raw, (size_t)total_nodes));
    k_leaves<<<bs_blocks(T.level_cnt[0], 128), 128, 0, st>>>(d_tris, d_ids2, n_tris, d_sorted, d_raw, T.level_cnt[0]);
    for (int l = 1; l < T.levels; ++l)
        k_level_up<<<bs_blocks(T.level_cnt[l], 128), 128, 0, st>>>(d_raw + T.level_off[l - 1], T.level_cnt[l - 1], d_raw + T.level_off[l], T.level_cnt[l]);
    k_finalize_nodes<<<bs_blocks(total_nodes, 128), 128, 0, st>>>(d_raw, total_nodes, d_nodes);
    bs_free(ctx, d_raw); bs_free(ctx, d_ids2);
    T.nodes = d_nodes; T.tris = d_sorted;
    bs_mark(ctx, "bvh_build_ms");
    if (vol->n_bricks) {
        const size_t warps = vol->n_bricks;
        size_t blocks = (warps + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
        const size_t max_blocks = (size_t)ctx->sm_count * 16 * 4;
        if (blocks > max_blocks) blocks = max_blocks;
        k_sign<<<(unsigned)blocks, 32 * WARPS_PER_BLOCK, 0, st>>>(T, vol->values, (unsigned long long*)vol->masks, vol->n_bricks, (const unsigned long long*)vol->keys, vol->voxel_size);
    }
    bs_mark(ctx, "sign_ms");
    bs_free(ctx, d_sorted); bs_free(ctx, d_nodes);
    BS_CUDA(ctx, cudaGetLastError());
    return BS_OK;
}
