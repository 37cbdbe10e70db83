// Volume::offset on sorted bricks: prune, fast-sweep Eikonal extension, shift (sm_100a).
// Replaces Volume::offset (src/voxel/volume/mod.rs:95-108) and FastSweeping (src/voxel/fast_sweep.rs:30-287,
// compute_distance :509-537).
//
// The reference runs ONE pass of 8 directional block Gauss-Seidel sweeps. Inside a sweep it pops leaves from a
// heap ordered lexicographically in the sweep direction, sweeps the 8^3 voxels of a leaf x-outer/z-inner in that
// direction (every update reads the six face neighbours as they are at that moment), then queues the next leaf
// along each axis whose exit face holds a value of the sweep's sign below the limit; everything popped in a
// sweep is queued again for the next one. That order is a topological order of "a leaf after its three upstream
// face neighbours and before its three downstream ones", and the same holds for voxels inside a leaf, so the
// result is reproduced exactly by a dataflow over the leaves (a leaf runs as soon as its upstream face neighbours have
// finished the sweep; all 8 sweeps in one persistent kernel, see k_sweep_all) and, inside a leaf, by a hyperplane
// wavefront over the voxels with equal +-x+-y+-z (22 steps). Dynamic queueing becomes a monotone per-brick flag set by
// the upstream leaf before it publishes its completion.
//
// Bricks the sweeps may create are pre-allocated empty (dilation of the pruned brick set); a push outside that
// set raises a flag and the whole operation is retried with a wider dilation.
#include "bs_common.cuh"
#include <cub/cub.cuh>
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <vector>

namespace {

typedef unsigned long long u64;
constexpr int TPB = 256;

__device__ __forceinline__ long long find_key(const u64* keys, size_t n, u64 k) {
    size_t lo = 0, hi = n;
    while (lo < hi) { size_t mid = (lo + hi) >> 1; if (keys[mid] < k) lo = mid + 1; else hi = mid; }
    return (lo < n && keys[lo] == k) ? (long long)lo : -1;
}

// remove_if(|v| > 2 vs) (volume/mod.rs:96; leaf_node/tree_node.rs remove_if): clears mask bits; flags non-empty bricks
__global__ void k_prune(const float* __restrict__ values, const u64* __restrict__ masks, size_t n, float thr, u64* out_masks, unsigned char* nonempty) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;  // one thread per mask word
    if (i >= n * 8) return;
    u64 m = masks[i], keep = 0;
    const float* v = values + i * 64;
    while (m) { const int b = __ffsll((long long)m) - 1; m &= m - 1; if (!(fabsf(v[b]) > thr)) keep |= 1ull << b; }
    out_masks[i] = keep;
    if (keep) nonempty[i >> 3] = 1;
}
// 27-neighbourhood dilation step of a brick key set
__global__ void k_dilate(const u64* __restrict__ keys, size_t n, u64* out, int* flags) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n * 27) return;
    const size_t b = i / 27; const int d = (int)(i % 27);
    int bx, by, bz; bs_key_brick(keys[b], bx, by, bz);
    bx += d / 9 - 1; by += (d / 3) % 3 - 1; bz += d % 3 - 1;
    if (bx < BS_BRICK_MIN || bx > BS_BRICK_MAX || by < BS_BRICK_MIN || by > BS_BRICK_MAX || bz < BS_BRICK_MIN || bz > BS_BRICK_MAX) { flags[0] = 1; out[i] = keys[b]; return; }
    out[i] = bs_brick_key(bx, by, bz);
}
struct IsSet { const unsigned char* f; __device__ bool operator()(const u64&) const { return true; } };
// scatter the pruned source bricks into the (sorted, dilated) working set and record the six face neighbours
__global__ void k_place(const u64* __restrict__ src_keys, const float* __restrict__ src_values, const u64* __restrict__ src_masks, const unsigned char* __restrict__ nonempty,
                        size_t n_src, const u64* __restrict__ keys, size_t n, float* values, u64* masks, u64* frozen, unsigned char* inq) {
    const size_t b = blockIdx.x;
    if (!nonempty[b]) return;
    const long long dst = find_key(keys, n, src_keys[b]);
    const unsigned t = threadIdx.x;
    values[dst * 512 + t] = src_values[b * 512 + t];
    if (t < 8) { masks[dst * 8 + t] = src_masks[b * 8 + t]; frozen[dst * 8 + t] = src_masks[b * 8 + t]; }
    if (t == 0) inq[dst] = 1;
}
__global__ void k_neighbours(const u64* __restrict__ keys, size_t n, int* nbr /*6 per brick: +x -x +y -y +z -z*/) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n * 6) return;
    const size_t b = i / 6; const int d = (int)(i % 6);
    int bx, by, bz; bs_key_brick(keys[b], bx, by, bz);
    const int s = (d & 1) ? -1 : 1;
    if (d < 2) bx += s; else if (d < 4) by += s; else bz += s;
    long long j = -1;
    if (bx >= BS_BRICK_MIN && bx <= BS_BRICK_MAX && by >= BS_BRICK_MIN && by <= BS_BRICK_MAX && bz >= BS_BRICK_MIN && bz <= BS_BRICK_MAX) j = find_key(keys, n, bs_brick_key(bx, by, bz));
    nbr[i] = (int)j;
}

// helpers/utils.rs:4-16 + fast_sweep.rs:509-537
__device__ __forceinline__ float compute_distance(float a1, float a2, float a3, float h) {
    float t;
    if (a1 > a3) { t = a1; a1 = a3; a3 = t; }
    if (a1 > a2) { t = a1; a1 = a2; a2 = t; }
    if (a2 > a3) { t = a2; a2 = a3; a3 = t; }
    const float s1 = xadd(a1, h);
    if (fabsf(s1) <= a2) return s1;
    const float a12 = xadd(a1, a2), hsq = xmul(h, h), two = xadd(hsq, hsq), d12 = xsub(a1, a2), d12s = xmul(d12, d12);
    const float s2 = xmul(xadd(a12, xsqrt(xsub(two, d12s))), 0.5f);
    if (fabsf(s2) <= a3) return s2;
    const float a123 = xadd(a12, a3), three = xadd(two, hsq), d13 = xsub(a1, a3), d13s = xmul(d13, d13), d23 = xsub(a2, a3), d23s = xmul(d23, d23);
    return xmul(xadd(a123, xsqrt(xsub(xsub(xsub(three, d12s), d13s), d23s))), __uint_as_float(0x3EAAAAABu) /* = 1.0f / 3.0f, correctly rounded */);
}

// wavefront index of every brick for the four (sx, sy, +z) patterns: w = sx bx + sy by + bz, biased to be >= 0
__global__ void k_wave_keys(const u64* __restrict__ keys, size_t n, unsigned* wk, unsigned* idx) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= 4 * n) return;
    const unsigned g = (unsigned)(i / n); const size_t b = i % n;
    int bx, by, bz; bs_key_brick(keys[b], bx, by, bz);
    const int w = ((g & 1) ? -bx : bx) + ((g & 2) ? -by : by) + bz;
    wk[i] = (unsigned)(w + 3 * (1 << 18)); idx[i] = (unsigned)b;
}

struct SweepParams {
    float* values; u64* masks; const u64* frozen; const int* nbr; unsigned char* inq; int* flags;
    const unsigned* order;   // [4][n]: bricks sorted by leaf wavefront for the four (sx, sy, +z) patterns
    unsigned n;              // bricks of the working set
    unsigned* done;          // [n] sweeps a brick has completed (task (sweep s, brick b) done  <=>  done[b] >= s + 1)
    float h, limit_abs; int sweep_neg;  // sweep sign: 1 = negative
};
__device__ __forceinline__ unsigned ld_done(const unsigned* p) { return *(const volatile unsigned*)p; }

// All 8 sweeps in ONE persistent kernel, as a dataflow over the tasks (sweep, leaf): task T = sweep * n + i takes the i-th
// leaf in the sweep's wavefront order (fast_sweep.rs:39-60: +++, -++, +-+, --+, ++-, -+-, +--, ---). The reference's order is
// a topological order of "a leaf after its three upstream face neighbours and before its three downstream ones"; a task may
// therefore start as soon as
//   * its three upstream neighbours have finished THIS sweep (their exit faces are what the stencil reads, and they are the
//     ones that queue this leaf), and
//   * the leaf itself and its three downstream neighbours have finished the PREVIOUS sweep (the stencil reads their faces as
//     that sweep left them; and nobody still reads this leaf's old values, because every neighbour is past that sweep),
// which is tracked by one counter per leaf (done[]). CTA c owns tasks c, c + G, c + 2G, ... of the global order: every wait is
// on a task with a smaller index, each CTA walks its tasks in increasing order and all G CTAs are resident, so the smallest
// unfinished task can always run -- no grid-wide barrier, no launch per wavefront (round 1: 1682 dependent launches), and
// consecutive sweeps overlap wherever the geometry allows. Data written by other CTAs is read with ld.global.cg (L2).
// One CTA = 64 threads = the (y, z) columns of a leaf; the brick lives in a 10^3 padded shared array (faces in the halo: no
// centre/face branches in the stencil), frozen bits sit in a register, the 22 voxel-wavefront steps touch shared memory only.
constexpr int PAD = 10, PAD2 = 100, PADN = 1000;
__global__ void __launch_bounds__(64) k_sweep_all(SweepParams P) {
    __shared__ float s_v[PADN];
    __shared__ unsigned char s_a[PADN];
    __shared__ unsigned char s_fz[64];
    const unsigned t = threadIdx.x;
    const unsigned ty = t >> 3, tz = t & 7;  // this thread's (y, z) column in brick coordinates (load / store phases)
    const unsigned long long n_tasks = 8ull * P.n;
    for (unsigned long long T = blockIdx.x; T < n_tasks; T += gridDim.x) {
        const int dir = (int)(T / P.n);          // bit0 = -x, bit1 = -y, bit2 = -z
        const unsigned i = (unsigned)(T % P.n);
        // sz = +1: pattern g = dir & 3 ascending; sz = -1: (sx, sy, -1) is the reverse of pattern (-sx, -sy, +1)
        const bool rev = dir & 4;
        const int g = rev ? ((~dir) & 3) : (dir & 3);
        const unsigned b = __ldg(P.order + (size_t)g * P.n + (rev ? P.n - 1 - i : i));
        const int sx = (dir & 1) ? -1 : 1, sy = (dir & 2) ? -1 : 1, sz = (dir & 4) ? -1 : 1;
        int nb[6];  // +x -x +y -y +z -z
#pragma unroll
        for (int d = 0; d < 6; ++d) nb[d] = __ldg(P.nbr + (size_t)b * 6 + d);
        // ---- dependencies, in two phases so that the loads of everything that is already final overlap the wait for the
        // upstream leaves (the chain of upstream waits is the critical path of the whole operation) ---------------------
        bool up[6];
#pragma unroll
        for (int d = 0; d < 6; ++d) { const int s = d < 2 ? sx : (d < 4 ? sy : sz); up[d] = (s > 0) == ((d & 1) != 0); }  // sweeping towards +: the - neighbour is upstream
        // phase 1: the leaf itself and its downstream neighbours are past the previous sweep (almost always true already)
        if (t < 7) {
            const int which = t < 6 ? nb[t] : (int)b;
            if (which >= 0 && !(t < 6 && up[t])) while (ld_done(P.done + which) < (unsigned)dir) __nanosleep(32);
        }
        __syncthreads();
        // a leaf queued in an earlier sweep stays queued: start its loads now (centre, masks, downstream faces)
        const bool early = __ldcg(P.inq + b) != 0;
        const float4* gv4 = reinterpret_cast<const float4*>(P.values + (size_t)b * 512);
        const unsigned sh = (ty << 3) | tz;
        const unsigned u = ty, v = tz;
        float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
        u64 mw[8], fw[8];
        float fv[6]; unsigned fa = 0;
        auto load_centre = [&]() {
            c0 = __ldcg(gv4 + t); c1 = __ldcg(gv4 + t + 64);
#pragma unroll
            for (int x = 0; x < 8; ++x) { mw[x] = __ldcg(P.masks + (size_t)b * 8 + x); fw[x] = __ldg(P.frozen + (size_t)b * 8 + x); }
        };
        auto load_face = [&](int d) {
            fv[d] = 0.f;
            if (nb[d] >= 0) {
                const unsigned c = (d & 1) ? 7u : 0u;  // the +x neighbour contributes its x = 0 face, the -x neighbour its x = 7 face
                const unsigned off = d < 2 ? ((c << 6) | (u << 3) | v) : (d < 4 ? ((u << 6) | (c << 3) | v) : ((u << 6) | (v << 3) | c));
                fa |= (unsigned)((__ldcg(P.masks + (size_t)nb[d] * 8 + (off >> 6)) >> (off & 63)) & 1) << d;
                fv[d] = __ldcg(P.values + (size_t)nb[d] * 512 + off);
            }
        };
        if (early) {
            load_centre();
#pragma unroll
            for (int d = 0; d < 6; ++d) if (!up[d]) load_face(d);
        }
        // phase 2: the upstream neighbours have finished THIS sweep
        if (t < 6 && up[t] && nb[t] >= 0) while (ld_done(P.done + nb[t]) < (unsigned)dir + 1u) __nanosleep(32);
        __syncthreads();
        const bool queued = early || __ldcg(P.inq + b) != 0;
        if (queued) {
            if (!early) {
                load_centre();
#pragma unroll
                for (int d = 0; d < 6; ++d) if (!up[d]) load_face(d);
            }
#pragma unroll
            for (int d = 0; d < 6; ++d) if (up[d]) load_face(d);
            unsigned abits = 0, fbits = 0;  // bit x = active / frozen flag of voxel (x, ty, tz)
#pragma unroll
            for (int x = 0; x < 8; ++x) { abits |= (unsigned)((mw[x] >> sh) & 1) << x; fbits |= (unsigned)((fw[x] >> sh) & 1) << x; }
            // fill the padded array: index (x+1)*100 + (y+1)*10 + (z+1)
            {
#pragma unroll
                for (int h = 0; h < 2; ++h) {  // centre: float4 q covers offsets 4q..4q+3 = (x, y, z0..z0+3)
                    const unsigned q = t + 64 * h, off = q * 4, x = off >> 6, y = (off >> 3) & 7, z = off & 7;
                    const float4 c = h ? c1 : c0;
                    float* d = s_v + (x + 1) * PAD2 + (y + 1) * PAD + (z + 1);
                    d[0] = c.x; d[1] = c.y; d[2] = c.z; d[3] = c.w;
                }
#pragma unroll
                for (int x = 0; x < 8; ++x) s_a[(x + 1) * PAD2 + (ty + 1) * PAD + (tz + 1)] = (abits >> x) & 1;
#pragma unroll
                for (int d = 0; d < 6; ++d) {
                    const unsigned c = (d & 1) ? 0u : 9u;  // +x face sits at padded x = 9, -x face at padded x = 0
                    const unsigned p = d < 2 ? (c * PAD2 + (u + 1) * PAD + (v + 1)) : (d < 4 ? ((u + 1) * PAD2 + c * PAD + (v + 1)) : ((u + 1) * PAD2 + (v + 1) * PAD + c));
                    s_v[p] = fv[d]; s_a[p] = (fa >> d) & 1;
                }
                s_fz[t] = (unsigned char)fbits;
            }
            __syncthreads();
            const int ly = t >> 3, lz = t & 7;  // sweep-local y, z of this thread's column
            const int y = sy > 0 ? ly : 7 - ly, z = sz > 0 ? lz : 7 - lz;
            const unsigned cfz = s_fz[((unsigned)y << 3) | (unsigned)z];  // frozen bits of column (y, z)
            const int pyz = (y + 1) * PAD + (z + 1);
            for (int step = 0; step < 22; ++step) {
                const int lx = step - ly - lz;
                if ((unsigned)lx < 8u) {
                    const int x = sx > 0 ? lx : 7 - lx;
                    if (!((cfz >> x) & 1)) {  // frozen voxels keep their value (:126-128)
                        const int p = (x + 1) * PAD2 + pyz;
                        // stencil.at for the six face neighbours (:130-146, :360-386)
                        const float nv[6] = {s_v[p + PAD2], s_v[p - PAD2], s_v[p + PAD], s_v[p - PAD], s_v[p + 1], s_v[p - 1]};
                        const bool na[6] = {s_a[p + PAD2] != 0, s_a[p - PAD2] != 0, s_a[p + PAD] != 0, s_a[p - PAD] != 0, s_a[p + 1] != 0, s_a[p - 1] != 0};
                        // option_min_by(+, -, cmp_abs): both present -> the smaller |v|, ties -> the + side
                        float d[3]; bool has[3];
#pragma unroll
                        for (int ax = 0; ax < 3; ++ax) {
                            const bool ap = na[2 * ax], an = na[2 * ax + 1];
                            has[ax] = ap || an;
                            d[ax] = (ap && an) ? ((fabsf(nv[2 * ax]) > fabsf(nv[2 * ax + 1])) ? nv[2 * ax + 1] : nv[2 * ax]) : (ap ? nv[2 * ax] : nv[2 * ax + 1]);
                        }
                        if (has[0] || has[1] || has[2]) {
                            const float first = has[0] ? d[0] : (has[1] ? d[1] : d[2]);
                            if ((int)(__float_as_uint(first) >> 31) == P.sweep_neg) {  // outward / inward (:148-156)
                                const float d1 = has[0] ? fabsf(d[0]) : FLT_MAX, d2 = has[1] ? fabsf(d[1]) : FLT_MAX, d3 = has[2] ? fabsf(d[2]) : FLT_MAX;
                                const float dn = compute_distance(d1, d2, d3, P.h);
                                if (!(dn > P.limit_abs)) {
                                    const float old = s_a[p] ? s_v[p] : FLT_MAX;
                                    if (dn < fabsf(old)) { s_v[p] = P.sweep_neg ? -fabsf(dn) : fabsf(dn); s_a[p] = 1; }  // set_sign (value/f32.rs:11-17)
                                }
                            }
                        }
                    }
                }
                __syncthreads();
            }
            // write back + queue the downstream leaves (:185-277)
            {
                float4* go4 = reinterpret_cast<float4*>(P.values + (size_t)b * 512);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const unsigned q = t + 64 * h, off = q * 4, x = off >> 6, yy = (off >> 3) & 7, zz = off & 7;
                    const float* d = s_v + (x + 1) * PAD2 + (yy + 1) * PAD + (zz + 1);
                    go4[q] = make_float4(d[0], d[1], d[2], d[3]);
                }
                // mask word x, bit (ty << 3 | tz) = t: each warp ballots its half of the word
                unsigned* gm = reinterpret_cast<unsigned*>(P.masks + (size_t)b * 8);
#pragma unroll
                for (int x = 0; x < 8; ++x) {
                    const unsigned bal = __ballot_sync(0xFFFFFFFFu, s_a[(x + 1) * PAD2 + (ty + 1) * PAD + (tz + 1)] != 0);
                    if ((t & 31) == 0) gm[2 * x + (t >> 5)] = bal;
                }
            }
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                const int s = ax == 0 ? sx : (ax == 1 ? sy : sz);
                const unsigned c = s > 0 ? 8u : 1u;  // exit face (padded coordinate); the reference's negative-direction index collapses to local 0 via leaf index masking
                const unsigned p = ax == 0 ? (c * PAD2 + (u + 1) * PAD + (v + 1)) : (ax == 1 ? ((u + 1) * PAD2 + c * PAD + (v + 1)) : ((u + 1) * PAD2 + (v + 1) * PAD + c));
                const bool q = s_a[p] && ((int)(__float_as_uint(s_v[p]) >> 31) == P.sweep_neg) && (fabsf(s_v[p]) < P.limit_abs);
                if (__syncthreads_or(q)) {
                    if (t == 0) {
                        const int n2 = s > 0 ? nb[2 * ax] : nb[2 * ax + 1];
                        if (n2 >= 0) P.inq[n2] = 1; else P.flags[0] = 1;  // outside the pre-allocated set: retry wider
                    }
                }
            }
        }
        // publish: everything this CTA wrote is visible before the counter moves (a leaf that was not queued wrote nothing)
        if (queued) __threadfence();
        __syncthreads();
        if (t == 0) *(volatile unsigned*)(P.done + b) = (unsigned)dir + 1u;
    }
}

__global__ void k_finish(float* values, const u64* __restrict__ masks, size_t n, float distance, unsigned char* nonempty) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;  // one thread per mask word
    if (i >= n * 8) return;
    u64 m = masks[i];
    if (m) nonempty[i >> 3] = 1;
    float* v = values + i * 64;
    while (m) { const int b = __ffsll((long long)m) - 1; m &= m - 1; v[b] = __fsub_rn(v[b], distance); }  // *v -= distance (volume/mod.rs:104)
}
__global__ void __launch_bounds__(512) k_compact(const u64* __restrict__ keys, const float* __restrict__ values, const u64* __restrict__ masks, const unsigned* __restrict__ rank,
                                                  const unsigned char* __restrict__ nonempty, u64* okeys, float* ovalues, u64* omasks) {
    const size_t b = blockIdx.x;
    if (!nonempty[b]) return;
    const size_t o = rank[b] - 1;
    const unsigned t = threadIdx.x;
    ovalues[o * 512 + t] = values[b * 512 + t];
    if (t < 8) omasks[o * 8 + t] = masks[b * 8 + t];
    if (t == 0) okeys[o] = keys[b];
}
__global__ void k_widen8(const unsigned char* in, unsigned* out, size_t n) { const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; if (i < n) out[i] = in[i]; }

struct KeyPass { __device__ bool operator()(const u64&) const { return true; } };

bs_status sort_unique(bs_context* ctx, u64* d_in, size_t n_in, u64** d_out, size_t* n_out) {
    cudaStream_t st = ctx->stream;
    u64* d_sorted = nullptr; size_t* d_n = nullptr; void* d_tmp = nullptr; size_t tmp = 0;
    BS_TRY(bs_alloc(ctx, &d_sorted, n_in)); BS_TRY(bs_alloc(ctx, d_out, n_in)); BS_TRY(bs_alloc(ctx, &d_n, 1));
    cub::DeviceRadixSort::SortKeys(nullptr, tmp, d_in, d_sorted, n_in, 0, 54, st);
    BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp));
    cub::DeviceRadixSort::SortKeys(d_tmp, tmp, d_in, d_sorted, n_in, 0, 54, st);
    bs_free(ctx, d_tmp); tmp = 0;
    cub::DeviceSelect::Unique(nullptr, tmp, d_sorted, *d_out, d_n, n_in, st);
    BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp));
    cub::DeviceSelect::Unique(d_tmp, tmp, d_sorted, *d_out, d_n, n_in, st);
    BS_CUDA(ctx, cudaMemcpyAsync(n_out, d_n, sizeof(size_t), cudaMemcpyDeviceToHost, st));
    BS_CUDA(ctx, cudaStreamSynchronize(st));
    bs_free(ctx, d_tmp); bs_free(ctx, d_sorted); bs_free(ctx, d_n);
    return BS_OK;
}

}  // namespace

bs_status bs_offset_impl(bs_volume* A, float distance, bs_volume** out) {
    bs_context* ctx = A->ctx;
    cudaStream_t st = ctx->stream;
    bs_marks_begin(ctx);
    const float vs = A->voxel_size;
    const size_t n_src = A->n_bricks;
    // limit = +-(|d| + vs + vs) (volume/mod.rs:98-99); active tiles (+-MAX) never survive the prune
    const float limit_abs = (fabsf(distance) + vs) + vs;
    const int sweep_neg = std::signbit(distance) ? 1 : 0;
    u64* d_pmasks = nullptr; unsigned char* d_nonempty = nullptr; int* d_flags = nullptr;
    BS_TRY(bs_alloc(ctx, &d_pmasks, n_src * 8)); BS_TRY(bs_alloc(ctx, &d_nonempty, n_src)); BS_TRY(bs_alloc(ctx, &d_flags, 1));
    BS_CUDA(ctx, cudaMemsetAsync(d_nonempty, 0, n_src ? n_src : 1, st));
    if (n_src) bs_count_launch(), k_prune<<<bs_blocks(n_src * 8, TPB), TPB, 0, st>>>(A->values, A->masks, n_src, vs * 2.0f, d_pmasks, d_nonempty);
    // keys of the non-empty pruned bricks
    u64* d_seed = nullptr; size_t n_seed = 0;
    {
        size_t* d_n = nullptr; void* d_tmp = nullptr; size_t tmp = 0;
        BS_TRY(bs_alloc(ctx, &d_seed, n_src)); BS_TRY(bs_alloc(ctx, &d_n, 1));
        cub::DeviceSelect::Flagged(nullptr, tmp, A->keys, d_nonempty, d_seed, d_n, n_src, st);
        BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp));
        cub::DeviceSelect::Flagged(d_tmp, tmp, A->keys, d_nonempty, d_seed, d_n, n_src, st);
        BS_CUDA(ctx, cudaMemcpyAsync(&n_seed, d_n, sizeof(size_t), cudaMemcpyDeviceToHost, st));
        BS_CUDA(ctx, cudaStreamSynchronize(st));
        bs_free(ctx, d_tmp); bs_free(ctx, d_n);
    }
    bs_mark(ctx, "offset_prune_ms");
    bs_volume* R = bs_volume_new(ctx, vs);
    if (n_seed == 0) {  // nothing within 2 voxels of the surface: the sweeps have nothing to extend
        bs_free(ctx, d_seed); bs_free(ctx, d_pmasks); bs_free(ctx, d_nonempty); bs_free(ctx, d_flags);
        bs_status s = bs_volume_alloc_bricks(R, 0);
        if (s != BS_OK) { bs_volume_free(R); return s; }
        bs_marks_end(ctx);
        *out = R;
        return BS_OK;
    }
    int K = (int)std::ceil(limit_abs / (8.0f * vs)) + 1;
    for (int attempt = 0;; ++attempt) {
        // --- working set: K dilation steps of the seed bricks ---------------------------------------------------
        u64* d_keys = nullptr; size_t n = n_seed;
        BS_TRY(bs_alloc(ctx, &d_keys, n));
        BS_CUDA(ctx, cudaMemcpyAsync(d_keys, d_seed, n * sizeof(u64), cudaMemcpyDeviceToDevice, st));
        BS_CUDA(ctx, cudaMemsetAsync(d_flags, 0, sizeof(int), st));
        for (int k = 0; k < K; ++k) {
            u64 *d_big = nullptr, *d_next = nullptr; size_t n_next = 0;
            BS_TRY(bs_alloc(ctx, &d_big, n * 27));
            bs_count_launch(), k_dilate<<<bs_blocks(n * 27, TPB), TPB, 0, st>>>(d_keys, n, d_big, d_flags);
            BS_TRY(sort_unique(ctx, d_big, n * 27, &d_next, &n_next));
            bs_free(ctx, d_big); bs_free(ctx, d_keys);
            d_keys = d_next; n = n_next;
        }
        float* d_values = nullptr; u64 *d_masks = nullptr, *d_frozen = nullptr; unsigned char* d_inq = nullptr; int* d_nbr = nullptr;
        BS_TRY(bs_alloc(ctx, &d_values, n * 512)); BS_TRY(bs_alloc(ctx, &d_masks, n * 8)); BS_TRY(bs_alloc(ctx, &d_frozen, n * 8));
        BS_TRY(bs_alloc(ctx, &d_inq, n)); BS_TRY(bs_alloc(ctx, &d_nbr, n * 6));
        BS_CUDA(ctx, cudaMemsetAsync(d_values, 0, n * 512 * sizeof(float), st));
        BS_CUDA(ctx, cudaMemsetAsync(d_masks, 0, n * 8 * sizeof(u64), st));
        BS_CUDA(ctx, cudaMemsetAsync(d_frozen, 0, n * 8 * sizeof(u64), st));
        BS_CUDA(ctx, cudaMemsetAsync(d_inq, 0, n, st));
        bs_count_launch(), k_place<<<(unsigned)n_src, 512, 0, st>>>(A->keys, A->values, d_pmasks, d_nonempty, n_src, d_keys, n, d_values, d_masks, d_frozen, d_inq);
        bs_count_launch(), k_neighbours<<<bs_blocks(n * 6, TPB), TPB, 0, st>>>(d_keys, n, d_nbr);
        // --- leaf wavefront order for the four axis-sign patterns (the other four are their reverses): bricks sorted by
        // w = +-bx +- by + bz on the device -----------------------------------------------------------------------
        unsigned *d_wk = nullptr, *d_wks = nullptr, *d_idx = nullptr, *d_order = nullptr;
        BS_TRY(bs_alloc(ctx, &d_wk, 4 * n)); BS_TRY(bs_alloc(ctx, &d_wks, 4 * n)); BS_TRY(bs_alloc(ctx, &d_idx, 4 * n)); BS_TRY(bs_alloc(ctx, &d_order, 4 * n));
        bs_count_launch(), k_wave_keys<<<bs_blocks(4 * n, TPB), TPB, 0, st>>>(d_keys, n, d_wk, d_idx);
        {
            void* d_tmp = nullptr; size_t tmp = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tmp, d_wk, d_wks, d_idx, d_order, (int)n, 0, 21, st);
            BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp));
            for (int g = 0; g < 4; ++g) cub::DeviceRadixSort::SortPairs(d_tmp, tmp, d_wk + g * n, d_wks + g * n, d_idx + g * n, d_order + g * n, (int)n, 0, 21, st);
            bs_free(ctx, d_tmp);
        }
        bs_free(ctx, d_wk); bs_free(ctx, d_wks); bs_free(ctx, d_idx);
        bs_mark(ctx, "offset_setup_ms");
        // --- 8 sweeps: one persistent kernel, every CTA resident -----------------------------------------------------------
        unsigned* d_done = nullptr;
        BS_TRY(bs_alloc(ctx, &d_done, n));
        BS_CUDA(ctx, cudaMemsetAsync(d_done, 0, n * sizeof(unsigned), st));
        SweepParams P;
        P.values = d_values; P.masks = d_masks; P.frozen = d_frozen; P.nbr = d_nbr; P.inq = d_inq; P.flags = d_flags;
        P.order = d_order; P.n = (unsigned)n; P.done = d_done;
        P.h = vs; P.limit_abs = limit_abs; P.sweep_neg = sweep_neg;
        int per_sm = 0;
        BS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_all, 64, 0));
        const unsigned grid = (unsigned)std::min<unsigned long long>(8ull * n, (unsigned long long)std::max(1, per_sm) * (unsigned long long)ctx->sm_count);
        bs_count_launch(), k_sweep_all<<<grid, 64, 0, st>>>(P);
        const size_t n_launch = 1;
        int flag = 0;
        BS_CUDA(ctx, cudaMemcpyAsync(&flag, d_flags, sizeof(int), cudaMemcpyDeviceToHost, st));
        BS_CUDA(ctx, cudaStreamSynchronize(st));
        bs_mark(ctx, "offset_sweep_ms");
        bs_status s = BS_OK;
        if (flag && attempt < 3) {  // pushed outside the working set (or index range hit): widen and redo
            bs_free(ctx, d_keys); bs_free(ctx, d_values); bs_free(ctx, d_masks); bs_free(ctx, d_frozen); bs_free(ctx, d_inq); bs_free(ctx, d_nbr); bs_free(ctx, d_order); bs_free(ctx, d_done);
            K += 2;
            continue;
        }
        if (flag) s = bs_fail(ctx, BS_ERR_RANGE, "offset: sweep left the supported index range");
        // --- remove_empty_branches + shift ----------------------------------------------------------------------
        unsigned char* d_ne = nullptr; unsigned *d_ne32 = nullptr, *d_rank = nullptr; void* d_tmp = nullptr; size_t tmp = 0; unsigned n_out = 0;
        if (s == BS_OK) s = bs_alloc(ctx, &d_ne, n);
        if (s == BS_OK) s = bs_alloc(ctx, &d_ne32, n);
        if (s == BS_OK) s = bs_alloc(ctx, &d_rank, n);
        if (s == BS_OK) {
            cudaMemsetAsync(d_ne, 0, n, st);
            bs_count_launch(), k_finish<<<bs_blocks(n * 8, TPB), TPB, 0, st>>>(d_values, d_masks, n, distance, d_ne);
            bs_count_launch(), k_widen8<<<bs_blocks(n, TPB), TPB, 0, st>>>(d_ne, d_ne32, n);
            cub::DeviceScan::InclusiveSum(nullptr, tmp, d_ne32, d_rank, n, st);
            s = bs_alloc(ctx, (char**)&d_tmp, tmp);
        }
        if (s == BS_OK) {
            cub::DeviceScan::InclusiveSum(d_tmp, tmp, d_ne32, d_rank, n, st);
            cudaMemcpyAsync(&n_out, d_rank + (n - 1), sizeof(unsigned), cudaMemcpyDeviceToHost, st);
            if (cudaStreamSynchronize(st) != cudaSuccess) s = bs_fail(ctx, BS_ERR_CUDA, "offset kernels failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
        if (s == BS_OK) s = bs_volume_alloc_bricks(R, n_out);
        if (s == BS_OK && n_out) bs_count_launch(), k_compact<<<(unsigned)n, 512, 0, st>>>(d_keys, d_values, d_masks, d_rank, d_ne, R->keys, R->values, R->masks);
        if (s == BS_OK && (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess)) s = bs_fail(ctx, BS_ERR_CUDA, "offset compaction failed");
        bs_free(ctx, d_tmp); bs_free(ctx, d_ne); bs_free(ctx, d_ne32); bs_free(ctx, d_rank);
        bs_free(ctx, d_keys); bs_free(ctx, d_values); bs_free(ctx, d_masks); bs_free(ctx, d_frozen); bs_free(ctx, d_inq); bs_free(ctx, d_nbr); bs_free(ctx, d_order); bs_free(ctx, d_done);
        bs_free(ctx, d_seed); bs_free(ctx, d_pmasks); bs_free(ctx, d_nonempty); bs_free(ctx, d_flags);
        if (s != BS_OK) { bs_volume_free(R); return s; }
        bs_mark(ctx, "offset_finish_ms");
        bs_marks_end(ctx);
        bs_stat_add(ctx, "n_work_bricks", (double)n);
        bs_stat_add(ctx, "n_out_bricks", (double)n_out);
        bs_stat_add(ctx, "n_sweep_launches", (double)n_launch);
        bs_stat_add(ctx, "dilation", (double)K);
        *out = R;
        return BS_OK;
    }
}
