// Sign propagation for closed meshes (sm_100a): the default sign path of MeshToVolume::compute_sings
// (src/voxel/mesh_to_volume.rs:198-281) whenever the input is a closed 2-cycle; open meshes keep the per-voxel
// winding-number traversal of bs_fwn.cu, which is the reference's semantics there.
//
// The reference evaluates the (approximate) winding number of EVERY active voxel and thresholds it at 0.2. On a closed
// mesh -- every directed edge a->b is matched by an edge b->a, i.e. the boundary of the triangle chain is zero -- the
// exact winding number is an integer that changes only across the surface: two lattice neighbours p, q have the same
// winding number unless a triangle meets the segment pq. k_block_edges (bs_convert.cu) marks exactly those lattice edges
// (conservatively, in fp64, from the original triangles); every other edge between two active voxels is a certified link.
//
// Steps: (1) closed? (2) connected components of the band under certified links: inside a brick a bit-parallel flood fill
// over the brick's 512-bit masks (8 lanes per brick, one 64-bit x-slab each; <= 8 components per brick, anything beyond
// becomes individually evaluated voxels), across brick faces a lock-free union-find over (brick, component) ids -- 8 ids
// per brick, not one per voxel; (3) ONE voxel per component is evaluated -- when there are only a few hundred (typically the
// inside shell, the outside shell and the voxels that lie on the surface within rounding) by one streaming pass over the
// triangles (k_sp_stream: groups of 32 consecutive triangles, far groups by their dipole expansion), else by the LBVH traversal of bs_fwn.cu on the representatives only; (4) every other
// voxel copies the sign of its component's representative.
//
// (1) has two implementations. Default: a 128-bit multiset fingerprint -- sum over directed edges of H(a, b) must equal
// the sum of H(b, a), H a 2 x 64-bit mix of the six coordinate words (-0 folded into +0; any non-finite coordinate means
// "not closed"): one streaming pass over the triangles instead of a hash table + 64-bit sort of 3 n edge keys; a false
// "closed" needs a 128-bit collision. BSHARK_CLOSED_CHECK=exact selects the sort-based test (vertices identified by exact
// coordinates like merge_points, every directed edge exactly once and its reverse exactly once);
// tests/test_gpu_signprop.py runs both on every test mesh.
#include "bs_common.cuh"
#include <cub/cub.cuh>
#include <cstdlib>
#include <cstring>

namespace {

typedef unsigned long long u64;
constexpr unsigned SP_EMPTY = 0xFFFFFFFFu;

// ---- (1a) closed? multiset fingerprint --------------------------------------------------------------------------------
__device__ __forceinline__ u64 sp_mix(u64 h) { h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33; return h; }
__device__ __forceinline__ void sp_edge_hash(const unsigned* a, const unsigned* b, u64& h0, u64& h1) {
    const u64 w0 = (u64)a[0] | ((u64)a[1] << 32), w1 = (u64)a[2] | ((u64)b[0] << 32), w2 = (u64)b[1] | ((u64)b[2] << 32);
    u64 x = sp_mix(w0 ^ 0x9E3779B97F4A7C15ULL); x = sp_mix(x ^ w1); x = sp_mix(x + w2);
    u64 y = sp_mix(w2 ^ 0xD6E8FEB86659FD93ULL); y = sp_mix(y + w0); y = sp_mix(y ^ (w1 * 0x9FB21C651E98DF25ULL));
    h0 = x; h1 = y;
}
__global__ void __launch_bounds__(256) k_sp_fingerprint(const float* __restrict__ tris, size_t n_tris, u64* out /*[4] fwd0 fwd1 rev0 rev1*/, int* bad) {
    u64 f0 = 0, f1 = 0, r0 = 0, r1 = 0; bool nonfinite = false;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n_tris; t += (size_t)gridDim.x * blockDim.x) {
        unsigned v[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const float f = tris[9 * t + i] + 0.0f;  // -0 -> +0
            if (!(fabsf(f) <= 3.0e38f)) nonfinite = true;
            v[i] = __float_as_uint(f);
        }
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            const unsigned* a = v + 3 * e; const unsigned* b = v + 3 * ((e + 1) % 3);
            u64 h0, h1;
            sp_edge_hash(a, b, h0, h1); f0 += h0; f1 += h1;
            sp_edge_hash(b, a, h0, h1); r0 += h0; r1 += h1;
        }
    }
    typedef cub::BlockReduce<u64, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    u64 s;
    s = BR(tmp).Sum(f0); if (threadIdx.x == 0) atomicAdd(out + 0, s); __syncthreads();
    s = BR(tmp).Sum(f1); if (threadIdx.x == 0) atomicAdd(out + 1, s); __syncthreads();
    s = BR(tmp).Sum(r0); if (threadIdx.x == 0) atomicAdd(out + 2, s); __syncthreads();
    s = BR(tmp).Sum(r1); if (threadIdx.x == 0) atomicAdd(out + 3, s);
    if (nonfinite) *bad = 1;
}

// ---- (1b) closed? exact test ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned sp_hash(float x, float y, float z) {
    const unsigned a = __float_as_uint(x + 0.0f), b = __float_as_uint(y + 0.0f), c = __float_as_uint(z + 0.0f);
    unsigned h = a * 73856093u ^ b * 19349663u ^ c * 83492791u;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
    return h;
}
// vertex id = slot of the vertex's coordinate class in an open-addressing table (one representative index per slot)
__global__ void k_sp_vertex_ids(const float* __restrict__ pts, size_t n, unsigned* table, unsigned mask, unsigned* vid) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    unsigned s = sp_hash(x, y, z) & mask;
    for (;;) {
        unsigned cur = table[s];
        if (cur == SP_EMPTY) { cur = atomicCAS(table + s, SP_EMPTY, (unsigned)i); if (cur == SP_EMPTY) break; }
        if (pts[3 * (size_t)cur] == x && pts[3 * (size_t)cur + 1] == y && pts[3 * (size_t)cur + 2] == z) break;
        s = (s + 1) & mask;  // (a NaN vertex never matches: it ends up alone, its edges unmatched -> "not closed")
    }
    vid[i] = s;
}
__global__ void k_sp_edges(const unsigned* __restrict__ vid, size_t n_tris, u64* keys, int* bad) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= n_tris) return;
    const unsigned a = vid[3 * t], b = vid[3 * t + 1], c = vid[3 * t + 2];
    if (a == b || b == c || c == a) *bad = 1;  // a triangle with a repeated vertex: be conservative
    keys[3 * t] = ((u64)a << 32) | b; keys[3 * t + 1] = ((u64)b << 32) | c; keys[3 * t + 2] = ((u64)c << 32) | a;
}
__global__ void k_sp_edges_check(const u64* __restrict__ keys /*sorted*/, size_t m, int* bad) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    const u64 k = keys[i];
    if (i + 1 < m && keys[i + 1] == k) { *bad = 1; return; }  // the same directed edge twice: non-manifold or inconsistent
    const u64 rev = (k << 32) | (k >> 32);
    size_t lo = 0, hi = m;
    while (lo < hi) { const size_t mid = (lo + hi) >> 1; if (keys[mid] < rev) lo = mid + 1; else hi = mid; }
    if (!(lo < m && keys[lo] == rev)) *bad = 1;               // boundary edge
}

// ---- (2) components -------------------------------------------------------------------------------------------------------
// Brick bit layout (bs_common.cuh): voxel (x, y, z) = bit (y << 3 | z) of word x. Face planes are 64-bit words:
// x faces: bit (y << 3 | z); y faces: bit (x << 3 | z); z faces: bit (x << 3 | y).
constexpr int SP_K = 8;  // components per brick
__device__ __forceinline__ u64 sp_gather8(u64 v) { return ((v & 0x0101010101010101ULL) * 0x0102040810204080ULL) >> 56; }  // bits 0, 8, .., 56 -> bits 0..7
__device__ __forceinline__ u64 sp_group_or(u64 v, unsigned gmask) {
    v |= __shfl_xor_sync(gmask, v, 1); v |= __shfl_xor_sync(gmask, v, 2); v |= __shfl_xor_sync(gmask, v, 4);
    return v;
}
// the six face planes (-x, +x, -y, +y, -z, +z) of a 512-bit set held one word per lane of an 8-lane group
__device__ __forceinline__ void sp_planes(u64 S, unsigned w, unsigned gmask, u64 out[6]) {
    out[0] = sp_group_or(w == 0 ? S : 0, gmask); out[1] = sp_group_or(w == 7 ? S : 0, gmask);
    out[2] = sp_group_or((S & 0xFF) << (8 * w), gmask); out[3] = sp_group_or((S >> 56) << (8 * w), gmask);
    out[4] = sp_group_or(sp_gather8(S) << (8 * w), gmask); out[5] = sp_group_or(sp_gather8(S >> 7) << (8 * w), gmask);
}
// 8 lanes per brick, lane w holds the x = w slab of every mask
__global__ void __launch_bounds__(256) k_sp_bricks(const u64* __restrict__ masks, const u64* __restrict__ blk, size_t n,
                                                   u64* comp /*[n][K][8]*/, u64* planes /*[n][K][6]*/, u64* face /*[n][9]: active -x +x -y +y -z +z, blocked +x +y +z*/,
                                                   u64* rest /*[n][8]*/, unsigned char* ncomp, unsigned short* first /*[n][K]*/, unsigned* par /*[n][K]*/) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t b = t >> 3;
    const unsigned w = (unsigned)(t & 7), gl = (threadIdx.x & 31) & ~7u, gmask = 0xFFu << gl;
    if (b >= n) return;  // whole groups leave together
    const u64 A = masks[b * 8 + w], BX = blk[b * 24 + w], BY = blk[b * 24 + 8 + w], BZ = blk[b * 24 + 16 + w];
    const u64 A_up = __shfl_down_sync(gmask, A, 1, 8);
    const u64 LZ = A & (A >> 1) & ~BZ & 0x7F7F7F7F7F7F7F7FULL;   // voxel linked to its +z neighbour
    const u64 LY = A & (A >> 8) & ~BY & 0x00FFFFFFFFFFFFFFULL;   // ... +y
    const u64 LX = w < 7 ? (A & A_up & ~BX) : 0;                 // ... +x (same bit of the next word)
    u64 LXdn = __shfl_up_sync(gmask, LX, 1, 8); if (w == 0) LXdn = 0;
    const u64 linked = LZ | (LZ << 1) | LY | (LY << 8) | LX | LXdn;
    // voxels without any link inside the brick and not on a face can have no link at all: evaluated on their own
    const u64 R = A & ~linked & ((w == 0 || w == 7) ? 0 : 0x007E7E7E7E7E7E00ULL);
    u64 U = A & ~R;
    unsigned k = 0;
    for (;;) {
        const unsigned nz = (__ballot_sync(gmask, U != 0) >> gl) & 0xFFu;
        if (nz == 0 || k == SP_K) break;
        const unsigned wl = __ffs(nz) - 1;
        u64 S = (w == wl) ? (U & (0 - U)) : 0;  // lowest voxel not yet in a component
        const unsigned f0 = (wl << 6) | (unsigned)(__ffsll((long long)__shfl_sync(gmask, U, gl + wl)) - 1);
        for (;;) {
            const u64 S0 = S;
#pragma unroll
            for (int r = 0; r < 3; ++r) S |= ((S & LZ) << 1) | ((S >> 1) & LZ) | ((S & LY) << 8) | ((S >> 8) & LY);
            u64 up = __shfl_up_sync(gmask, S & LX, 1, 8); if (w == 0) up = 0;
            const u64 dn = __shfl_down_sync(gmask, S, 1, 8) & LX;  // LX is 0 on lane 7
            S |= up | dn;
            if (!__any_sync(gmask, S != S0)) break;
        }
        comp[(b * SP_K + k) * 8 + w] = S;
        U &= ~S;
        u64 pl[6];
        sp_planes(S, w, gmask, pl);
        if (w < 6) planes[(b * SP_K + k) * 6 + w] = pl[w];
        if (w == 0) first[b * SP_K + k] = (unsigned short)f0;
        ++k;
    }
    rest[b * 8 + w] = R | U;  // (U != 0 only when the brick holds more than SP_K components)
    par[b * SP_K + w] = (unsigned)(b * SP_K + w);
    if (w == 0) ncomp[b] = (unsigned char)k;
    u64 pa[6], pb[6];
    sp_planes(A, w, gmask, pa);
    if (w < 6) face[b * 9 + w] = pa[w];
    // blocked bits of the edges leaving through +x (word 7 of BX), +y (y = 7 rows of BY), +z (z = 7 bits of BZ)
    sp_planes(BX, w, gmask, pb); const u64 bx = pb[1];
    sp_planes(BY, w, gmask, pb); const u64 by = pb[3];
    sp_planes(BZ, w, gmask, pb); const u64 bz = pb[5];
    if (w == 0) { face[b * 9 + 6] = bx; face[b * 9 + 7] = by; face[b * 9 + 8] = bz; }
}
// union-find over (brick, component) ids: the larger root is hooked under the smaller one by atomicCAS; finds halve the
// path as they go (a halving write only ever replaces a parent by one of its ancestors). Loads bypass L1.
__device__ __forceinline__ unsigned sp_find_g(unsigned* par, unsigned x) {
    volatile unsigned* vp = par;
    unsigned p = vp[x];
    while (p != x) {
        const unsigned g = vp[p];
        if (g != p) vp[x] = g;
        x = p; p = g;
    }
    return x;
}
__device__ __forceinline__ void sp_union_g(unsigned* par, unsigned a, unsigned b) {
    for (;;) {
        a = sp_find_g(par, a); b = sp_find_g(par, b);
        if (a == b) return;
        if (a > b) { const unsigned t = a; a = b; b = t; }
        if (atomicCAS(par + b, b, a) == b) return;
    }
}
__device__ __forceinline__ long long sp_find_key(const u64* keys, size_t n, u64 k) {
    size_t lo = 0, hi = n;
    while (lo < hi) { const size_t mid = (lo + hi) >> 1; if (keys[mid] < k) lo = mid + 1; else hi = mid; }
    return (lo < n && keys[lo] == k) ? (long long)lo : -1;
}
// one thread per (brick, axis): certified links across the brick's +axis face join components of the two bricks
__global__ void __launch_bounds__(256) k_sp_faces(const u64* __restrict__ keys, size_t n, const u64* __restrict__ planes, const u64* __restrict__ face,
                                                  const unsigned char* __restrict__ ncomp, unsigned* par) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= n * 3) return;
    const size_t b = t / 3; const int a = (int)(t % 3);
    const unsigned nc = ncomp[b];
    if (nc == 0) return;
    int bx, by, bz; bs_key_brick(keys[b], bx, by, bz);
    bx += a == 0; by += a == 1; bz += a == 2;
    if (bx > BS_BRICK_MAX || by > BS_BRICK_MAX || bz > BS_BRICK_MAX) return;
    const long long nb = sp_find_key(keys, n, bs_brick_key(bx, by, bz));
    if (nb < 0) return;
    const u64 L = face[b * 9 + 2 * a + 1] & face[(size_t)nb * 9 + 2 * a] & ~face[b * 9 + 6 + a];
    if (L == 0) return;
    const unsigned nn = ncomp[nb];
    for (unsigned k = 0; k < nc; ++k) {
        const u64 pk = planes[(b * SP_K + k) * 6 + 2 * a + 1] & L;
        if (pk == 0) continue;
        for (unsigned j = 0; j < nn; ++j)
            if (pk & planes[((size_t)nb * SP_K + j) * 6 + 2 * a]) sp_union_g(par, (unsigned)(b * SP_K + k), (unsigned)((size_t)nb * SP_K + j));
    }
}
// 8 lanes per brick, lane k = component k: flatten (read-only finds: a halving write by another thread could land after
// this thread's own store and leave a non-root behind) and build the mask of the voxels to evaluate: the leftover voxels
// plus the first voxel of every component that is its class's root.
__global__ void __launch_bounds__(256) k_sp_flatten(size_t n, unsigned* par, const unsigned char* __restrict__ ncomp, const unsigned short* __restrict__ first,
                                                    const u64* __restrict__ rest, u64* seed_masks, unsigned* n_chunks, int per_chunk, u64* n_seeds) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t b = t >> 3;
    const unsigned k = (unsigned)(t & 7), gl = (threadIdx.x & 31) & ~7u, gmask = 0xFFu << gl;
    if (b >= n) return;
    const unsigned id = (unsigned)(b * SP_K + k);
    bool root = false; unsigned off = 0;
    if (k < ncomp[b]) {
        unsigned r = id, p; volatile unsigned* vp = par;
        while ((p = vp[r]) != r) r = p;
        par[id] = r;
        root = r == id; off = first[id];
    }
    u64 word = rest[b * 8 + k];  // lane k doubles as the owner of word k
#pragma unroll
    for (int j = 0; j < SP_K; ++j) {
        const unsigned oj = __shfl_sync(gmask, root ? off : 0xFFFFu, gl + j);
        if (oj != 0xFFFFu && (oj >> 6) == k) word |= 1ull << (oj & 63);
    }
    seed_masks[b * 8 + k] = word;
    unsigned c = (unsigned)__popcll(word);
    c += __shfl_xor_sync(gmask, c, 1); c += __shfl_xor_sync(gmask, c, 2); c += __shfl_xor_sync(gmask, c, 4);
    if (k == 0) { n_chunks[b] = (c + per_chunk - 1) / per_chunk; if (c) atomicAdd(n_seeds, (u64)c); }
}

// ---- (3) up to SP_BRUTE_MAX representatives: one streaming pass over the triangles, no tree ------------------------------------
// A warp takes 32 consecutive triangles of the input (in most meshes a spatially compact strip), reduces them to the
// reference's node moments (aabb_tree.rs:723-801: area, area-weighted centre and normal, order-2 tensor, bounding radius about
// the centre) and then serves the representatives 32 at a time, one per lane: far ones (|p - p~| > 2.5 r, the reference
// accepts at 2 r) get the order-1 + order-2 expansion (:667-669, 803-816), near ones the exact solid angles of the 32
// triangles, one triangle per lane (:582-615). Input order only affects speed: a scattered group has a large radius and
// is evaluated exactly.
constexpr int SP_BRUTE_MAX = 512, SP_NC = SP_BRUTE_MAX / 32;
__global__ void k_sp_collect(const u64* __restrict__ seed_masks, size_t n_words, unsigned* list /*voxel ids brick*512+off*/, unsigned* count, unsigned cap) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n_words) return;
    u64 m = seed_masks[i];
    while (m) {
        const int bit = __ffsll((long long)m) - 1; m &= m - 1;
        const unsigned slot = atomicAdd(count, 1u);
        if (slot < cap) list[slot] = (unsigned)(i * 64 + bit);
    }
}
__device__ __forceinline__ float sp_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}
__global__ void __launch_bounds__(256) k_sp_stream(const float* __restrict__ tris, size_t n_tris, const unsigned* __restrict__ list, unsigned n_list,
                                                   const u64* __restrict__ keys, float vs, double* wn /*[n_list]*/) {
    __shared__ float s_q[SP_BRUTE_MAX][3];
    __shared__ double s_acc[SP_BRUTE_MAX];
    for (unsigned i = threadIdx.x; i < SP_BRUTE_MAX; i += blockDim.x) {
        s_acc[i] = 0.0;
        if (i < n_list) {
            const unsigned g = list[i], off = g & 511;
            int bx, by, bz; bs_key_brick(keys[g >> 9], bx, by, bz);
            s_q[i][0] = __fmul_rn((float)((bx << 3) + (int)(off >> 6)), vs);
            s_q[i][1] = __fmul_rn((float)((by << 3) + (int)((off >> 3) & 7)), vs);
            s_q[i][2] = __fmul_rn((float)((bz << 3) + (int)(off & 7)), vs);
        } else { s_q[i][0] = s_q[i][1] = s_q[i][2] = 0.f; }
    }
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    const unsigned nc = (n_list + 31) / 32;
    float acc[SP_NC];
#pragma unroll
    for (int c = 0; c < SP_NC; ++c) acc[c] = 0.f;
    const size_t n_groups = (n_tris + 31) / 32, warp0 = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t g = warp0; g < n_groups; g += n_warps) {
        const size_t t = g * 32 + lane;
        float v[9];
        if (t < n_tris) { const float* p = tris + 9 * t; for (int i = 0; i < 9; ++i) v[i] = p[i]; } else { for (int i = 0; i < 9; ++i) v[i] = 0.f; }
        const float e1x = v[3] - v[0], e1y = v[4] - v[1], e1z = v[5] - v[2], e2x = v[6] - v[0], e2y = v[7] - v[1], e2z = v[8] - v[2];
        float hx = 0.5f * (e1y * e2z - e1z * e2y), hy = 0.5f * (e1z * e2x - e1x * e2z), hz = 0.5f * (e1x * e2y - e1y * e2x);  // area * normal
        float area = sqrtf(hx * hx + hy * hy + hz * hz);
        if (!(area > 0.f) || !(area < 3.0e38f)) { area = 0.f; hx = hy = hz = 0.f; }  // degenerate triangles are skipped (:757-759)
        const float cx = (v[0] + v[3] + v[6]) * (1.0f / 3.0f), cy = (v[1] + v[4] + v[7]) * (1.0f / 3.0f), cz = (v[2] + v[5] + v[8]) * (1.0f / 3.0f);
        const float A = sp_warp_sum(area);
        if (!(A > 0.f)) continue;  // warp-uniform
        const float iA = 1.0f / A;
        const float px = sp_warp_sum(area * cx) * iA, py = sp_warp_sum(area * cy) * iA, pz = sp_warp_sum(area * cz) * iA;  // p~
        const float ox = sp_warp_sum(hx), oy = sp_warp_sum(hy), oz = sp_warp_sum(hz);                                       // order 1
        // order 2 about p~: M[col][row] = sum (c[row] - p~[row]) * h[col]; only tr M and the symmetric parts enter r^T M r
        const float dx = cx - px, dy = cy - py, dz = cz - pz;
        const float m00 = sp_warp_sum(dx * hx), m11 = sp_warp_sum(dy * hy), m22 = sp_warp_sum(dz * hz);
        const float m01 = sp_warp_sum(dy * hx + dx * hy), m02 = sp_warp_sum(dz * hx + dx * hz), m12 = sp_warp_sum(dz * hy + dy * hz);
        float r2max = 0.f;
        if (t < n_tris) for (int k = 0; k < 3; ++k) { const float ax = v[3 * k] - px, ay = v[3 * k + 1] - py, az = v[3 * k + 2] - pz; r2max = fmaxf(r2max, ax * ax + ay * ay + az * az); }
#pragma unroll
        for (int o = 16; o; o >>= 1) r2max = fmaxf(r2max, __shfl_xor_sync(0xFFFFFFFFu, r2max, o));
        const float far2 = 6.25f * r2max;  // (2.5 r)^2
#pragma unroll
        for (int c = 0; c < SP_NC; ++c) {
            if ((unsigned)c >= nc) break;  // uniform
            const unsigned si = c * 32 + lane;
            const float qx = s_q[si][0], qy = s_q[si][1], qz = s_q[si][2];
            const float rx = px - qx, ry = py - qy, rz = pz - qz;
            const float r2 = rx * rx + ry * ry + rz * rz;
            const bool valid = si < n_list, far = valid && r2 > far2;
            if (far) {
                const float inv_r = rsqrtf(r2), ir2 = inv_r * inv_r, k = 0.07957747154594767f * inv_r * ir2;
                const float d = ox * rx + oy * ry + oz * rz + (m00 + m11 + m22);
                const float rMr = rx * (m00 * rx + m01 * ry + m02 * rz) + ry * (m11 * ry + m12 * rz) + m22 * rz * rz;
                acc[c] += k * (d - 3.0f * ir2 * rMr);
            }
            unsigned near = __ballot_sync(0xFFFFFFFFu, valid && !far);
            while (near) {  // exact: lane = triangle, one representative at a time
                const int j = __ffs(near) - 1; near &= near - 1;
                const float sx = s_q[c * 32 + j][0], sy = s_q[c * 32 + j][1], sz = s_q[c * 32 + j][2];
                const float ax = v[0] - sx, ay = v[1] - sy, az = v[2] - sz, bx = v[3] - sx, by = v[4] - sy, bz = v[5] - sz, ccx = v[6] - sx, ccy = v[7] - sy, ccz = v[8] - sz;
                const float la = sqrtf(ax * ax + ay * ay + az * az), lb = sqrtf(bx * bx + by * by + bz * bz), lc = sqrtf(ccx * ccx + ccy * ccy + ccz * ccz);
                float w = 0.f;
                if (la != 0.f && lb != 0.f && lc != 0.f) {
                    const float det = ax * (by * ccz - bz * ccy) + ay * (bz * ccx - bx * ccz) + az * (bx * ccy - by * ccx);
                    const float den = la * lb * lc + (ax * bx + ay * by + az * bz) * lc + (ax * ccx + ay * ccy + az * ccz) * lb + (bx * ccx + by * ccy + bz * ccz) * la;
                    if (det != 0.f) w = atan2f(det, den) * (2.0f * 0.07957747154594767f);
                }
                w = sp_warp_sum(w);
                if ((int)lane == j) acc[c] += w;
            }
        }
    }
#pragma unroll
    for (int c = 0; c < SP_NC; ++c) if ((unsigned)c < nc && acc[c] != 0.f) atomicAdd(&s_acc[c * 32 + lane], (double)acc[c]);
    __syncthreads();
    for (unsigned i = threadIdx.x; i < n_list; i += blockDim.x) if (s_acc[i] != 0.0) atomicAdd(wn + i, s_acc[i]);
}
__global__ void k_sp_brute_apply(float* values, const unsigned* __restrict__ list, unsigned n_list, const double* __restrict__ wn) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_list) return;
    const float d = values[list[i]];
    values[list[i]] = ((float)wn[i] < 0.2f) ? copysignf(d, 1.0f) : copysignf(d, -1.0f);  // mesh_to_volume.rs:266-271
}

// ---- (4) every voxel of a component takes the sign its class's representative got -----------------------------------------
// one warp per brick
__global__ void __launch_bounds__(256) k_sp_broadcast(float* values, size_t n, const u64* __restrict__ comp, const unsigned char* __restrict__ ncomp,
                                                      const unsigned short* __restrict__ first, const unsigned* __restrict__ par) {
    const size_t b = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31;
    if (b >= n) return;
    const unsigned nc = ncomp[b];
    if (nc == 0) return;
    bool neg = false;
    if (lane < nc) {
        const unsigned r = par[b * SP_K + lane];  // flattened: the class's root component
        neg = (__float_as_uint(__ldcg(values + (size_t)(r / SP_K) * 512 + first[r])) >> 31) != 0;
    }
    const unsigned negm = __ballot_sync(0xFFFFFFFFu, neg);
    u64 ALL = 0, NEG = 0;
    if (lane < 8)
        for (unsigned k = 0; k < nc; ++k) { const u64 m = comp[(b * SP_K + k) * 8 + lane]; ALL |= m; if ((negm >> k) & 1u) NEG |= m; }
    float* bv = values + b * 512;
#pragma unroll 4
    for (int r = 0; r < 16; ++r) {
        const u64 all = __shfl_sync(0xFFFFFFFFu, ALL, r >> 1), ng = __shfl_sync(0xFFFFFFFFu, NEG, r >> 1);
        const unsigned sh = (r & 1) * 32 + lane;
        if ((all >> sh) & 1ull) {
            const float v = bv[r * 32 + lane];
            bv[r * 32 + lane] = ((ng >> sh) & 1ull) ? -fabsf(v) : fabsf(v);
        }
    }
}

}  // namespace

static bs_status mesh_closed_exact(bs_context* ctx, const float* d_tris, size_t n_tris, bool* closed) {
    cudaStream_t st = ctx->stream;
    const size_t nv = n_tris * 3;
    if (nv >= 0x7FFFFFFFull) return BS_OK;
    size_t cap = 1024; while (cap < 2 * nv) cap <<= 1;
    unsigned *table = nullptr, *vid = nullptr; u64 *k1 = nullptr, *k2 = nullptr; int* d_bad = nullptr; void* d_tmp = nullptr; size_t tmp = 0;
    BS_TRY(bs_alloc(ctx, &table, cap)); BS_TRY(bs_alloc(ctx, &vid, nv)); BS_TRY(bs_alloc(ctx, &k1, nv)); BS_TRY(bs_alloc(ctx, &k2, nv)); BS_TRY(bs_alloc(ctx, &d_bad, 1));
    BS_CUDA(ctx, cudaMemsetAsync(table, 0xFF, cap * sizeof(unsigned), st));
    BS_CUDA(ctx, cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    bs_count_launch(), k_sp_vertex_ids<<<bs_blocks(nv, 256), 256, 0, st>>>(d_tris, nv, table, (unsigned)(cap - 1), vid);
    bs_count_launch(), k_sp_edges<<<bs_blocks(n_tris, 256), 256, 0, st>>>(vid, n_tris, k1, d_bad);
    cub::DeviceRadixSort::SortKeys(nullptr, tmp, k1, k2, (int)nv, 0, 64, st);
    BS_TRY(bs_alloc(ctx, (char**)&d_tmp, tmp));
    cub::DeviceRadixSort::SortKeys(d_tmp, tmp, k1, k2, (int)nv, 0, 64, st);
    bs_count_launch(), k_sp_edges_check<<<bs_blocks(nv, 256), 256, 0, st>>>(k2, nv, d_bad);
    int bad = 1;
    BS_TRY(bs_fetch(ctx, &bad, d_bad, sizeof(int)));
    BS_TRY(bs_sync(ctx));
    bs_free(ctx, d_tmp); bs_free(ctx, table); bs_free(ctx, vid); bs_free(ctx, k1); bs_free(ctx, k2); bs_free(ctx, d_bad);
    *closed = bad == 0;
    return BS_OK;
}

// Starts the closedness test on the stream; bs_mesh_closed_finish reads the verdict (after the caller's next sync)
bs_status bs_mesh_closed_begin(bs_context* ctx, const float* d_tris, size_t n_tris, bs_closed_check* chk) {
    cudaStream_t st = ctx->stream;
    chk->exact = false; chk->closed = false; chk->d_sums = nullptr; chk->d_bad = nullptr; chk->pending = false;
    if (n_tris == 0) return BS_OK;
    const char* e = getenv("BSHARK_CLOSED_CHECK");
    if (e && strcmp(e, "exact") == 0) { chk->exact = true; return mesh_closed_exact(ctx, d_tris, n_tris, &chk->closed); }
    BS_TRY(bs_alloc(ctx, &chk->d_sums, 4)); BS_TRY(bs_alloc(ctx, &chk->d_bad, 1));
    BS_CUDA(ctx, cudaMemsetAsync(chk->d_sums, 0, 4 * sizeof(u64), st));
    BS_CUDA(ctx, cudaMemsetAsync(chk->d_bad, 0, sizeof(int), st));
    const unsigned grid = (unsigned)std::min<size_t>(bs_blocks(n_tris, 256), (size_t)ctx->sm_count * 16);
    bs_count_launch(), k_sp_fingerprint<<<grid, 256, 0, st>>>(d_tris, n_tris, chk->d_sums, chk->d_bad);
    BS_TRY(bs_fetch(ctx, chk->h_sums, chk->d_sums, 4 * sizeof(u64)));
    BS_TRY(bs_fetch(ctx, &chk->h_bad, chk->d_bad, sizeof(int)));
    chk->pending = true;
    return BS_OK;
}
// call after a stream synchronisation that follows bs_mesh_closed_begin
bool bs_mesh_closed_finish(bs_context* ctx, bs_closed_check* chk) {
    if (chk->pending) {
        chk->closed = chk->h_bad == 0 && chk->h_sums[0] == chk->h_sums[2] && chk->h_sums[1] == chk->h_sums[3];
        bs_free(ctx, chk->d_sums); bs_free(ctx, chk->d_bad);
        chk->d_sums = nullptr; chk->d_bad = nullptr; chk->pending = false;
    }
    return chk->closed;
}

// Components of the band voxels of `vol` (masks = active bits, d_blk = blocked lattice edges from k_block_edges).
// On return C holds the per-brick component tables, C->seed the masks of the voxels to evaluate, d_nchunks[b] the work
// items (of per_chunk representatives) of brick b and *d_nseeds their total. C->ok = false (nothing allocated): the
// volume has too many bricks for 32-bit component ids and the caller keeps the per-voxel path.
bs_status bs_sign_components_impl(bs_context* ctx, const bs_volume* vol, const unsigned long long* d_blk, bs_sign_components* C, unsigned* d_nchunks, int per_chunk, unsigned long long* d_nseeds) {
    cudaStream_t st = ctx->stream;
    const size_t n = vol->n_bricks;
    memset(C, 0, sizeof(*C));
    if (n == 0 || n * SP_K >= 0xFFFFFFFFull || !d_blk) return BS_OK;
    BS_TRY(bs_alloc(ctx, &C->comp, n * SP_K * 8)); BS_TRY(bs_alloc(ctx, &C->planes, n * SP_K * 6)); BS_TRY(bs_alloc(ctx, &C->face, n * 9));
    BS_TRY(bs_alloc(ctx, &C->rest, n * 8)); BS_TRY(bs_alloc(ctx, &C->ncomp, n)); BS_TRY(bs_alloc(ctx, &C->first, n * SP_K));
    BS_TRY(bs_alloc(ctx, &C->par, n * SP_K)); BS_TRY(bs_alloc(ctx, &C->seed, n * 8));
    bs_count_launch(), k_sp_bricks<<<bs_blocks(n * 8, 256), 256, 0, st>>>(vol->masks, d_blk, n, C->comp, C->planes, C->face, C->rest, C->ncomp, C->first, C->par);
    bs_count_launch(), k_sp_faces<<<bs_blocks(n * 3, 256), 256, 0, st>>>(vol->keys, n, C->planes, C->face, C->ncomp, C->par);
    bs_count_launch(), k_sp_flatten<<<bs_blocks(n * 8, 256), 256, 0, st>>>(n, C->par, C->ncomp, C->first, C->rest, C->seed, d_nchunks, per_chunk, d_nseeds);
    BS_CUDA(ctx, cudaGetLastError());
    C->ok = true;
    return BS_OK;
}
void bs_sign_components_free(bs_context* ctx, bs_sign_components* C) {
    bs_free(ctx, C->comp); bs_free(ctx, C->planes); bs_free(ctx, C->face); bs_free(ctx, C->rest); bs_free(ctx, C->ncomp); bs_free(ctx, C->first); bs_free(ctx, C->par); bs_free(ctx, C->seed);
    memset(C, 0, sizeof(*C));
}
int bs_sign_brute_max() { return SP_BRUTE_MAX; }
// evaluates the (<= SP_BRUTE_MAX) representatives in one streaming pass over all triangles and sets their signs
bs_status bs_sign_brute_impl(bs_context* ctx, const float* d_tris, size_t n_tris, bs_volume* vol, const bs_sign_components* C, unsigned n_seeds) {
    cudaStream_t st = ctx->stream;
    if (n_seeds == 0) return BS_OK;
    unsigned *d_list = nullptr, *d_count = nullptr; double* d_wn = nullptr;
    BS_TRY(bs_alloc(ctx, &d_list, (size_t)SP_BRUTE_MAX)); BS_TRY(bs_alloc(ctx, &d_count, 1)); BS_TRY(bs_alloc(ctx, &d_wn, (size_t)SP_BRUTE_MAX));
    BS_CUDA(ctx, cudaMemsetAsync(d_count, 0, sizeof(unsigned), st));
    BS_CUDA(ctx, cudaMemsetAsync(d_wn, 0, SP_BRUTE_MAX * sizeof(double), st));
    bs_count_launch(), k_sp_collect<<<bs_blocks(vol->n_bricks * 8, 256), 256, 0, st>>>(C->seed, vol->n_bricks * 8, d_list, d_count, (unsigned)SP_BRUTE_MAX);
    const unsigned grid = (unsigned)std::min<size_t>(bs_blocks((n_tris + 31) / 32 * 32, 256), (size_t)ctx->sm_count * 6);
    bs_count_launch(), k_sp_stream<<<grid, 256, 0, st>>>(d_tris, n_tris, d_list, n_seeds, vol->keys, vol->voxel_size, d_wn);
    bs_count_launch(), k_sp_brute_apply<<<bs_blocks(SP_BRUTE_MAX, 256), 256, 0, st>>>(vol->values, d_list, n_seeds, d_wn);
    bs_free(ctx, d_list); bs_free(ctx, d_count); bs_free(ctx, d_wn);
    BS_CUDA(ctx, cudaGetLastError());
    return BS_OK;
}
bs_status bs_sign_broadcast_impl(bs_context* ctx, bs_volume* vol, const bs_sign_components* C) {
    const size_t n = vol->n_bricks;
    if (n) bs_count_launch(), k_sp_broadcast<<<bs_blocks(n * 32, 256), 256, 0, ctx->stream>>>(vol->values, n, C->comp, C->ncomp, C->first, C->par);
    BS_CUDA(ctx, cudaGetLastError());
    return BS_OK;
}
