// Volume construction from explicit voxels and from analytic primitives (sm_100a).
// Replaces Volume::from_fn (src/voxel/volume/mod.rs:40-72) and VolumeBuilder::{sphere, cuboid, iwp}
// (src/voxel/volume/builder.rs:21-76). The reference walks the dense index box serially and inserts voxel by
// voxel into the tree; here a brick-granular pass over the same box keeps the bricks that own at least one
// voxel with |f| <= (band + 1) * voxel_size and a second pass writes them in sorted-key order.
#include "bs_common.cuh"
#include <cub/cub.cuh>
#include <cfloat>

namespace {

constexpr int TPB = 256;

// ---- from explicit voxels -------------------------------------------------------------------------------------
__global__ void k_voxel_keys(const int32_t* __restrict__ ijk, size_t m, unsigned long long* comp, unsigned* idx, int* flags) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int x = ijk[3 * i], y = ijk[3 * i + 1], z = ijk[3 * i + 2];
    const int bx = x >> 3, by = y >> 3, bz = z >> 3;
    if (bx < BS_BRICK_MIN || bx > BS_BRICK_MAX || by < BS_BRICK_MIN || by > BS_BRICK_MAX || bz < BS_BRICK_MIN || bz > BS_BRICK_MAX) { flags[0] = 1; comp[i] = 0; idx[i] = (unsigned)i; return; }
    const unsigned off = ((x & 7) << 6) | ((y & 7) << 3) | (z & 7);
    comp[i] = (bs_brick_key(bx, by, bz) << 9) | off;
    idx[i] = (unsigned)i;
}
__global__ void k_brick_heads(const unsigned long long* __restrict__ comp, size_t m, unsigned* head) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    head[i] = (i == 0 || (comp[i] >> 9) != (comp[i - 1] >> 9)) ? 1u : 0u;
}
__global__ void k_scatter_voxels(const unsigned long long* __restrict__ comp, const unsigned* __restrict__ idx, const unsigned* __restrict__ rank /*inclusive scan of head*/,
                                 const float* __restrict__ val, size_t m, unsigned long long* keys, float* values, unsigned long long* masks) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    const unsigned long long c = comp[i];
    const unsigned b = rank[i] - 1, off = (unsigned)(c & 511);
    if (i == 0 || (c >> 9) != (comp[i - 1] >> 9)) keys[b] = c >> 9;
    if (i + 1 < m && comp[i + 1] == c) return;  // duplicates: the LAST insert wins (stable sort keeps input order)
    values[(size_t)b * 512 + off] = val[idx[i]];
    atomicOr(&masks[(size_t)b * 8 + (off >> 6)], 1ull << (off & 63));
}

// ---- analytic primitives --------------------------------------------------------------------------------------
struct Prim {
    int kind;         // 0 sphere, 1 cuboid, 2 iwp
    float p[7];       // sphere: radius, origin.xyz ; cuboid: min.xyz max.xyz ; iwp: min.xyz max.xyz cell_size
    float vs, nbw;    // voxel size, (band + 1) * voxel size
    int lo[3], hi[3]; // inclusive voxel index box of from_fn
    int blo[3], bdim[3];
};

__device__ __forceinline__ float box_sq_dist(const float* mn, const float* mx, float x, float y, float z) {  // box3.rs:95-111
    float sq = 0.f;
    const float v[3] = {x, y, z};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (v[i] < mn[i]) { const float d = xsub(mn[i], v[i]); sq = xadd(sq, xmul(d, d)); }
        if (v[i] > mx[i]) { const float d = xsub(v[i], mx[i]); sq = xadd(sq, xmul(d, d)); }
    }
    return sq;
}
__device__ __forceinline__ bool box_contains(const float* mn, const float* mx, float x, float y, float z) {
    return x >= mn[0] && x <= mx[0] && y >= mn[1] && y <= mx[1] && z >= mn[2] && z <= mx[2];
}
__device__ float prim_eval(const Prim& P, int ix, int iy, int iz) {
    const float x = xmul((float)ix, P.vs), y = xmul((float)iy, P.vs), z = xmul((float)iz, P.vs);
    if (P.kind == 0) {  // builder.rs:21-30: (p - origin).norm() - radius
        const f3 d = xsub(f3{x, y, z}, f3{P.p[1], P.p[2], P.p[3]});
        return xsub(xsqrt(xnorm2(d)), P.p[0]);
    }
    if (P.kind == 1) {  // builder.rs:32-51
        const float* mn = P.p; const float* mx = P.p + 3;
        if (box_contains(mn, mx, x, y, z)) {
            float m = fminf(xsub(x, mn[0]), xsub(mx[0], x));
            m = fminf(m, xsub(y, mn[1])); m = fminf(m, xsub(mx[1], y));
            m = fminf(m, xsub(z, mn[2])); m = fminf(m, xsub(mx[2], z));
            return -m;
        }
        return xsqrt(box_sq_dist(mn, mx, x, y, z));
    }
    // builder.rs:54-76 (cosf here vs the host libm in the reference: values agree to a few ulp)
    const float* mn = P.p; const float* mx = P.p + 3;
    const float cell = P.p[6], inv = xdiv(1.0f, cell);
    const float imn[3] = {xsub(mn[0], P.vs), xsub(mn[1], P.vs), xsub(mn[2], P.vs)}, imx[3] = {xadd(mx[0], P.vs), xadd(mx[1], P.vs), xadd(mx[2], P.vs)};
    if (!box_contains(mn, mx, x, y, z)) return xsqrt(box_sq_dist(imn, imx, x, y, z));
    const float cx = cosf(xmul(x, inv)), cy = cosf(xmul(y, inv)), cz = cosf(xmul(z, inv));
    const float s = xadd(xadd(cx, cy), cz);
    const float pr = xadd(xadd(xmul(cx, cy), xmul(cy, cz)), xmul(cz, cx));
    const float v = -xsub(xsub(s, xmul(0.51f, pr)), 1.0f);
    return xmul(v, cell);
}

// pass 1: one CTA of 512 threads per brick of the dense brick box -> keep flag
__global__ void __launch_bounds__(512) k_prim_flags(Prim P, unsigned char* keep) {
    const int bi = blockIdx.x;
    const int bx = P.blo[0] + bi / (P.bdim[1] * P.bdim[2]), by = P.blo[1] + (bi / P.bdim[2]) % P.bdim[1], bz = P.blo[2] + bi % P.bdim[2];
    const int t = threadIdx.x;
    const int ix = (bx << 3) + (t >> 6), iy = (by << 3) + ((t >> 3) & 7), iz = (bz << 3) + (t & 7);
    bool in = ix >= P.lo[0] && ix <= P.hi[0] && iy >= P.lo[1] && iy <= P.hi[1] && iz >= P.lo[2] && iz <= P.hi[2];
    bool k = false;
    if (in) { const float v = prim_eval(P, ix, iy, iz); k = !(fabsf(v) > P.nbw); }
    const int any = __syncthreads_or(k);
    if (t == 0) keep[bi] = (unsigned char)(any != 0);
}
__global__ void k_prim_keys(Prim P, const unsigned char* __restrict__ keep, size_t n, unsigned long long* keys) {
    size_t bi = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (bi >= n) return;
    const int bx = P.blo[0] + (int)(bi / ((size_t)P.bdim[1] * P.bdim[2])), by = P.blo[1] + (int)((bi / P.bdim[2]) % P.bdim[1]), bz = P.blo[2] + (int)(bi % P.bdim[2]);
    keys[bi] = keep[bi] ? bs_brick_key(bx, by, bz) : BS_KEY_INVALID;
}
struct NotInvalid { __device__ bool operator()(unsigned long long k) const { return k != BS_KEY_INVALID; } };
// pass 2: one CTA per kept brick (sorted): values + mask
__global__ void __launch_bounds__(512) k_prim_fill(Prim P, const unsigned long long* __restrict__ keys, float* values, unsigned long long* masks) {
    const size_t b = blockIdx.x;
    int bx, by, bz; bs_key_brick(keys[b], bx, by, bz);
    const int t = threadIdx.x;
    const int ix = (bx << 3) + (t >> 6), iy = (by << 3) + ((t >> 3) & 7), iz = (bz << 3) + (t & 7);
    const bool in = ix >= P.lo[0] && ix <= P.hi[0] && iy >= P.lo[1] && iy <= P.hi[1] && iz >= P.lo[2] && iz <= P.hi[2];
    float v = 0.f; bool k = false;
    if (in) { v = prim_eval(P, ix, iy, iz); k = !(fabsf(v) > P.nbw); }
    values[b * 512 + t] = k ? v : 0.f;
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, k);
    __shared__ unsigned s_bal[16];
    if ((t & 31) == 0) s_bal[t >> 5] = bal;
    __syncthreads();
    if (t < 8) masks[b * 8 + t] = (unsigned long long)s_bal[2 * t] | ((unsigned long long)s_bal[2 * t + 1] << 32);
}

}  // namespace

bs_status bs_from_voxels_impl(bs_context* ctx, const int32_t* d_ijk, const float* d_values, size_t m, float voxel_size, bs_volume** out) {
    cudaStream_t st = ctx->stream;
    if (m >= (1ull << 32)) return bs_fail(ctx, BS_ERR_RANGE, "too many voxels");
    unsigned long long *d_comp = nullptr, *d_comp2 = nullptr; unsigned *d_idx = nullptr, *d_idx2 = nullptr, *d_head = nullptr, *d_rank = nullptr; int* d_flags = nullptr;
    BS_TRY(bs_alloc(ctx, &d_comp, m)); BS_TRY(bs_alloc(ctx, &d_comp2, m)); BS_TRY(bs_alloc(ctx, &d_idx, m)); BS_TRY(bs_alloc(ctx, &d_idx2, m));
    BS_TRY(bs_alloc(ctx, &d_head, m)); BS_TRY(bs_alloc(ctx, &d_rank, m)); BS_TRY(bs_alloc(ctx, &d_flags, 1));
    BS_CUDA(ctx, cudaMemsetAsync(d_flags, 0, sizeof(int), st));
    bs_count_launch(), k_voxel_keys<<<bs_blocks(m, TPB), TPB, 0, st>>>(d_ijk, m, d_comp, d_idx, d_flags);
    void* d_tmp = nullptr; size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_comp, d_comp2, d_idx, d_idx2, m, 0, 63, st);
    BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp_bytes));
    cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_comp, d_comp2, d_idx, d_idx2, m, 0, 63, st);
    bs_free(ctx, d_tmp);
    bs_count_launch(), k_brick_heads<<<bs_blocks(m, TPB), TPB, 0, st>>>(d_comp2, m, d_head);
    tmp_bytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, d_head, d_rank, m, st);
    BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp_bytes));
    cub::DeviceScan::InclusiveSum(d_tmp, tmp_bytes, d_head, d_rank, m, st);
    unsigned n_bricks = 0; int flag = 0;
    BS_CUDA(ctx, cudaMemcpyAsync(&n_bricks, d_rank + (m - 1), sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    BS_CUDA(ctx, cudaMemcpyAsync(&flag, d_flags, sizeof(int), cudaMemcpyDeviceToHost, st));
    BS_CUDA(ctx, cudaStreamSynchronize(st));
    bs_free(ctx, d_tmp);
    bs_status s = BS_OK;
    bs_volume* v = nullptr;
    if (flag) s = bs_fail(ctx, BS_ERR_RANGE, "voxel index outside [-2^20, 2^20)");
    if (s == BS_OK) {
        v = bs_volume_new(ctx, voxel_size);
        s = bs_volume_alloc_bricks(v, n_bricks);
    }
    if (s == BS_OK) {
        cudaMemsetAsync(v->values, 0, (size_t)n_bricks * 512 * sizeof(float), st);
        cudaMemsetAsync(v->masks, 0, (size_t)n_bricks * 8 * sizeof(unsigned long long), st);
        bs_count_launch(), k_scatter_voxels<<<bs_blocks(m, TPB), TPB, 0, st>>>(d_comp2, d_idx2, d_rank, d_values, m, v->keys, v->values, v->masks);
        if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) s = bs_fail(ctx, BS_ERR_CUDA, "from_voxels kernels failed");
    }
    bs_free(ctx, d_comp); bs_free(ctx, d_comp2); bs_free(ctx, d_idx); bs_free(ctx, d_idx2); bs_free(ctx, d_head); bs_free(ctx, d_rank); bs_free(ctx, d_flags);
    if (s != BS_OK) { if (v) bs_volume_free(v); return s; }
    *out = v;
    return BS_OK;
}

bs_status bs_builder_impl(bs_context* ctx, int kind, float voxel_size, const float* p, bs_volume** out) {
    cudaStream_t st = ctx->stream;
    *out = nullptr;
    if (!(voxel_size > 0.0f)) return bs_fail(ctx, BS_ERR_INVALID, "voxel_size must be > 0");
    Prim P;
    P.kind = kind; P.vs = voxel_size;
    for (int i = 0; i < 7; ++i) P.p[i] = p[i];
    float mn[3], mx[3]; int band;
    if (kind == 0) {  // builder.rs:22-26
        band = 1;
        const float off = p[0] + (float)band * voxel_size;
        for (int d = 0; d < 3; ++d) { mn[d] = p[1 + d] + (-off); mx[d] = p[1 + d] + off; }
    } else if (kind == 1) {  // builder.rs:33-37
        band = 1;
        const float off = (float)band * voxel_size;
        for (int d = 0; d < 3; ++d) { mn[d] = p[d] + (-off); mx[d] = p[3 + d] + off; }
    } else {  // builder.rs:57-62: iwp bbox = [min - vs, max + vs], band 2
        band = 2;
        for (int d = 0; d < 3; ++d) { mn[d] = p[d] + (-voxel_size); mx[d] = p[3 + d] + voxel_size; }
    }
    P.nbw = (float)(band + 1) * voxel_size;  // volume/mod.rs:49
    for (int d = 0; d < 3; ++d) {
        const float lo = floorf(mn[d] / voxel_size), hi = ceilf(mx[d] / voxel_size);  // volume/mod.rs:50-51
        if (!(lo > -1.0e6f && hi < 1.0e6f)) return bs_fail(ctx, BS_ERR_RANGE, "voxel index outside [-2^20, 2^20)");
        P.lo[d] = (int)lo; P.hi[d] = (int)hi;
        P.blo[d] = P.lo[d] >> 3; P.bdim[d] = (P.hi[d] >> 3) - P.blo[d] + 1;
        if (P.bdim[d] <= 0) return bs_volume_empty(ctx, voxel_size, out);
    }
    const size_t n_dense = (size_t)P.bdim[0] * P.bdim[1] * P.bdim[2];
    if (n_dense > (1ull << 31) - 1) return bs_fail(ctx, BS_ERR_RANGE, "primitive box too large");
    unsigned char* d_keep = nullptr; unsigned long long *d_keys = nullptr, *d_sel = nullptr; size_t* d_nsel = nullptr;
    BS_TRY(bs_alloc(ctx, &d_keep, n_dense)); BS_TRY(bs_alloc(ctx, &d_keys, n_dense)); BS_TRY(bs_alloc(ctx, &d_sel, n_dense)); BS_TRY(bs_alloc(ctx, &d_nsel, 1));
    bs_count_launch(), k_prim_flags<<<(unsigned)n_dense, 512, 0, st>>>(P, d_keep);
    bs_count_launch(), k_prim_keys<<<bs_blocks(n_dense, TPB), TPB, 0, st>>>(P, d_keep, n_dense, d_keys);
    void* d_tmp = nullptr; size_t tmp_bytes = 0;
    cub::DeviceSelect::If(nullptr, tmp_bytes, d_keys, d_sel, d_nsel, n_dense, NotInvalid(), st);
    BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp_bytes));
    cub::DeviceSelect::If(d_tmp, tmp_bytes, d_keys, d_sel, d_nsel, n_dense, NotInvalid(), st);
    size_t n = 0;
    BS_CUDA(ctx, cudaMemcpyAsync(&n, d_nsel, sizeof(size_t), cudaMemcpyDeviceToHost, st));
    BS_CUDA(ctx, cudaStreamSynchronize(st));
    bs_free(ctx, d_tmp); bs_free(ctx, d_keep); bs_free(ctx, d_keys); bs_free(ctx, d_nsel);
    bs_volume* v = bs_volume_new(ctx, voxel_size);
    bs_status s = bs_volume_alloc_bricks(v, n);
    if (s == BS_OK && n) {
        tmp_bytes = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, d_sel, v->keys, n, 0, 54, st);
        s = bs_alloc(ctx, (char**)&d_tmp, tmp_bytes);
        if (s == BS_OK) {
            cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, d_sel, v->keys, n, 0, 54, st);
            bs_count_launch(), k_prim_fill<<<(unsigned)n, 512, 0, st>>>(P, v->keys, v->values, v->masks);
            bs_free(ctx, d_tmp);
            if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) s = bs_fail(ctx, BS_ERR_CUDA, "builder kernels failed");
        }
    }
    bs_free(ctx, d_sel);
    if (s != BS_OK) { bs_volume_free(v); return s; }
    *out = v;
    return BS_OK;
}
