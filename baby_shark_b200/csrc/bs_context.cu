// Context, volume handles and the extern "C" surface of libbshark_cuda (include/bshark.h).
#include "bs_common.cuh"
#include "mc33_tables.h"
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <algorithm>

bs_status bs_fail(bs_context* ctx, bs_status st, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    if (ctx) { ctx->err = buf; ctx->failed_epoch = ctx->epoch; ctx->fail_pending = true; ctx->pending.clear(); ctx->ctrl_used = 0; }  // (the failing call unwinds: its pending read-backs point into dead frames)
    return st;
}

namespace {
constexpr size_t CTRL_WORDS = 1 << 14;
__global__ void k_fetch(const unsigned* __restrict__ src, unsigned* dst, unsigned n) {
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}
}  // namespace
bs_status bs_fetch(bs_context* ctx, void* host_dst, const void* d_src, size_t bytes) {
    if (bytes == 0) return BS_OK;
    if ((bytes & 3) || ((uintptr_t)d_src & 3) || bytes > CTRL_WORDS * 2) return bs_fail(ctx, BS_ERR_INVALID, "bs_fetch: %zu bytes", bytes);
    if (!ctx->h_ctrl) {
        BS_CUDA(ctx, cudaHostAlloc((void**)&ctx->h_ctrl, CTRL_WORDS * sizeof(unsigned), cudaHostAllocMapped));
        BS_CUDA(ctx, cudaHostGetDevicePointer((void**)&ctx->d_ctrl, ctx->h_ctrl, 0));
    }
    const size_t words = bytes / 4;
    if (ctx->ctrl_used + words > CTRL_WORDS) BS_TRY(bs_sync(ctx));  // scratch full: deliver what is pending first
    k_fetch<<<1, 64, 0, ctx->stream>>>((const unsigned*)d_src, ctx->d_ctrl + ctx->ctrl_used, (unsigned)words);
    ctx->pending.push_back({host_dst, ctx->ctrl_used, bytes});
    ctx->ctrl_used += words;
    return BS_OK;
}
bs_status bs_sync(bs_context* ctx) {
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return bs_fail(ctx, BS_ERR_CUDA, "stream synchronize: %s", cudaGetErrorString(e));
    for (const auto& p : ctx->pending) memcpy(p.dst, ctx->h_ctrl + p.off, p.bytes);
    ctx->pending.clear(); ctx->ctrl_used = 0;
    return BS_OK;
}

void bs_marks_begin(bs_context* ctx) {
    for (auto& m : ctx->marks) cudaEventDestroy(m.second);
    ctx->marks.clear(); ctx->stats.clear();
    bs_mark(ctx, "begin");
}
void bs_mark(bs_context* ctx, const char* name) {
    cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, ctx->stream);
    ctx->marks.push_back({name, e});
}
void bs_marks_end(bs_context* ctx) {
    bs_sync(ctx);
    for (size_t i = 1; i < ctx->marks.size(); ++i) {
        float ms = 0.f; cudaEventElapsedTime(&ms, ctx->marks[i - 1].second, ctx->marks[i].second);
        ctx->stats.push_back({ctx->marks[i].first, (double)ms});
    }
    if (ctx->marks.size() > 1) {
        float ms = 0.f; cudaEventElapsedTime(&ms, ctx->marks.front().second, ctx->marks.back().second);
        ctx->stats.push_back({"total_ms", (double)ms});
    }
}
void bs_stat_add(bs_context* ctx, const char* name, double v) { ctx->stats.push_back({name, v}); }

bs_volume* bs_volume_new(bs_context* ctx, float voxel_size) {
    bs_volume* v = new bs_volume();
    v->ctx = ctx; v->voxel_size = voxel_size;
    ctx->live_volumes.insert(v);
    return v;
}
bs_status bs_volume_alloc_bricks(bs_volume* v, size_t n) {
    v->n_bricks = n; v->n_owned = n;
    BS_TRY(bs_alloc(v->ctx, &v->keys, n));
    BS_TRY(bs_alloc(v->ctx, &v->values, n * 512));
    BS_TRY(bs_alloc(v->ctx, &v->masks, n * 8));
    return BS_OK;
}

static const int8_t h_mc33[MC33_BLOB_SIZE] = MC33_BLOB_INIT;

template <class T> static bs_status dup_array(bs_context* c, T** dst, const T* src, size_t n) {
    BS_TRY(bs_alloc(c, dst, n));
    if (n) BS_CUDA(c, cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyDeviceToDevice, c->stream));
    return BS_OK;
}
template <class T> static bs_status to_host(bs_context* c, T** dst, const T* src, size_t n) {
    *dst = nullptr;
    BS_CUDA(c, cudaMallocHost((void**)dst, (n ? n : 1) * sizeof(T)));
    if (n) BS_CUDA(c, cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
    return BS_OK;
}

unsigned long long g_bs_launches = 0;

// one source, up to 16 destinations (peers' buffers mapped over NVLink, own buffer included): each value is read once and
// stored `world` times. A slice starts at a multiple of 9 floats, not of 16 bytes, and source and destination are out of
// phase: the body is cut into DESTINATION-aligned float4 stores (NVLink likes wide writes) fed by four scalar loads, the
// ragged head and tail go out as single floats.
struct PushDst { float* p[16]; };
__global__ void __launch_bounds__(256) k_push_out_verts(const float* __restrict__ src, PushDst D, int world, size_t n, unsigned head /*floats before the first 16 B boundary of the destinations*/) {
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    const size_t h = head < n ? head : n, body4 = (n - h) / 4, tail0 = h + body4 * 4;
    for (size_t q = tid; q < body4; q += stride) {
        const float* s = src + h + q * 4;
        const float4 v = make_float4(s[0], s[1], s[2], s[3]);
#pragma unroll 4
        for (int d = 0; d < world; ++d) reinterpret_cast<float4*>(D.p[d] + h)[q] = v;
    }
    if (tid < h) { const float v = src[tid]; for (int d = 0; d < world; ++d) D.p[d][tid] = v; }
    if (tid < n - tail0) { const float v = src[tail0 + tid]; for (int d = 0; d < world; ++d) D.p[d][tail0 + tid] = v; }
}

// ---- block cache ------------------------------------------------------------------------------------------------------
// cudaMallocAsync on the default pool showed sporadic 100-500 ms stalls in steady state (pool growth / remapping when
// the mix of sizes fragments it). The library's allocation pattern is the same every call, so freed blocks are kept
// by rounded size (granularity 1/8 of the power of two below the request, >= 512 B: <= 12.5 % slack) and reused.
static size_t bs_round_size(size_t bytes) {
    if (bytes <= 512) return 512;
    size_t p = 512;
    while ((p << 1) <= bytes) p <<= 1;
    const size_t g = std::max<size_t>(512, p >> 3);
    return (bytes + g - 1) / g * g;
}
void bs_cache_release(bs_context* ctx) {
    cudaStreamSynchronize(ctx->stream);
    for (auto& kv : ctx->cache_free) { cudaFree(kv.second); ctx->cache_total_bytes -= kv.first; }
    ctx->cache_free.clear(); ctx->cache_free_bytes = 0;
}
void bs_op_begin(bs_context* ctx) {
    if (ctx->fail_pending) {  // the failed call returned early: whatever it still holds goes back to the cache (nothing of it was handed out)
        std::vector<void*> dead;
        for (auto& kv : ctx->cache_live) if (kv.second.epoch == ctx->failed_epoch) dead.push_back(kv.first);
        for (void* p : dead) bs_raw_free(ctx, p);
        ctx->fail_pending = false;
    }
    ++ctx->epoch;
}
bs_status bs_raw_alloc(bs_context* ctx, size_t bytes, void** out) {
    const size_t sz = bs_round_size(bytes);
    auto it = ctx->cache_free.find(sz);
    void* p = nullptr;
    if (it != ctx->cache_free.end()) { p = it->second; ctx->cache_free.erase(it); ctx->cache_free_bytes -= sz; }
    else {
        cudaError_t e = cudaMalloc(&p, sz);
        if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); bs_cache_release(ctx); e = cudaMalloc(&p, sz); }
        if (e != cudaSuccess) { cudaGetLastError(); return bs_fail(ctx, BS_ERR_CUDA, "device allocation of %zu bytes failed: %s", sz, cudaGetErrorString(e)); }
        ctx->cache_total_bytes += sz;
    }
    ctx->cache_live[p] = bs_context::LiveBlock{sz, ctx->epoch};
    *out = p;
    return BS_OK;
}
void bs_raw_free(bs_context* ctx, void* p) {
    auto it = ctx->cache_live.find(p);
    if (it == ctx->cache_live.end()) return;
    const size_t sz = it->second.size;
    ctx->cache_live.erase(it);
    ctx->cache_free.emplace(sz, p); ctx->cache_free_bytes += sz;
    if (ctx->cache_free_bytes > ((size_t)48 << 30)) bs_cache_release(ctx);  // varying workloads: do not hoard HBM
}

template <class T> static bs_status to_pinned(bs_context* c, const T* d_src, size_t n, T** out) {
    *out = nullptr;
    T* h = nullptr;
    BS_CUDA(c, cudaMallocHost((void**)&h, (n ? n : 1) * sizeof(T)));
    if (n && cudaMemcpyAsync(h, d_src, n * sizeof(T), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { cudaFreeHost(h); return bs_fail(c, BS_ERR_CUDA, "device to host copy failed"); }
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { cudaFreeHost(h); return bs_fail(c, BS_ERR_CUDA, "device to host copy failed"); }
    *out = h;
    return BS_OK;
}

extern "C" {

bs_status bs_context_create(int device, bs_context** out) {
    if (!out) return BS_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return BS_ERR_NO_DEVICE;
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) return BS_ERR_NO_DEVICE; }
    if (device >= count) return BS_ERR_INVALID;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return BS_ERR_NO_DEVICE;
    if (prop.major != 10) return BS_ERR_NO_DEVICE;  // kernels are built for sm_100a only
    bs_context* ctx = new bs_context();
    ctx->device = device; ctx->sm_count = prop.multiProcessorCount;
    if (const char* e = getenv("BSHARK_SIGN_PROPAGATION")) ctx->sign_propagation = atoi(e);  // A/B runs: 0 = per-voxel signs everywhere
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx; return BS_ERR_CUDA;
    }
    if (cudaMalloc((void**)&ctx->d_err, sizeof(unsigned)) != cudaSuccess || cudaMemset(ctx->d_err, 0, sizeof(unsigned)) != cudaSuccess ||
        cudaMalloc((void**)&ctx->d_mc33, MC33_BLOB_SIZE) != cudaSuccess ||
        cudaMemcpy(ctx->d_mc33, h_mc33, MC33_BLOB_SIZE, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaStreamDestroy(ctx->stream); delete ctx; return BS_ERR_CUDA;
    }
    *out = ctx;
    return BS_OK;
}
void bs_context_destroy(bs_context* ctx) {
    if (!ctx) return;
    {
    BS_ENTER(ctx);
    cudaStreamSynchronize(ctx->stream);
    // volumes the caller still holds lose their context (their device memory goes with it): bs_volume_free on such a handle
    // only deletes the host object, every other call on it is BS_ERR_INVALID
    for (bs_volume* v : ctx->live_volumes) { v->ctx = nullptr; v->keys = nullptr; v->values = nullptr; v->masks = nullptr; v->owned = nullptr; v->tile8_keys = nullptr; v->tile8_values = nullptr; v->tile128_keys = nullptr; v->tile128_values = nullptr; v->n_bricks = 0; }
    ctx->live_volumes.clear();
    for (auto& m : ctx->marks) cudaEventDestroy(m.second);
    if (ctx->d_out_verts) cudaFree(ctx->d_out_verts);
    if (ctx->d_out_alt) cudaFree(ctx->d_out_alt);
    if (ctx->d_mc_carry) cudaFree(ctx->d_mc_carry);
    if (ctx->copy_stream) { cudaStreamDestroy(ctx->copy_stream); for (int i = 0; i < 2; ++i) { cudaEventDestroy(ctx->ev_done[i]); cudaEventDestroy(ctx->ev_copied[i]); } }
    bs_cache_release(ctx);
    for (auto& kv : ctx->cache_live) cudaFree(kv.first);  // volumes the caller never freed
    cudaFree(ctx->d_mc33); cudaFree(ctx->d_err);
    if (ctx->h_ctrl) cudaFreeHost(ctx->h_ctrl);
    cudaStreamDestroy(ctx->stream);
    }
    delete ctx;
}
unsigned long long bs_kernel_launch_count(void) { return g_bs_launches; }
const char* bs_last_error(const bs_context* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int bs_context_device(const bs_context* ctx) { return ctx ? ctx->device : -1; }
void* bs_context_stream(const bs_context* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
bs_status bs_context_set_flag(bs_context* ctx, int flag, int value) {
    if (!ctx) return BS_ERR_INVALID;
    if (flag == BS_FLAG_COUNT_WORK) { ctx->count_work = value; return BS_OK; }
    if (flag == BS_FLAG_SIGN_PROPAGATION) { ctx->sign_propagation = value; return BS_OK; }
    return BS_ERR_INVALID;
}
bs_status bs_context_copy_out_verts_device(bs_context* ctx, float* d_dst, size_t n_floats) {
    if (!ctx || (!d_dst && n_floats)) return BS_ERR_INVALID;
    BS_ENTER(ctx);  // before any bs_fail: a failure belongs to THIS call's epoch
    if (n_floats > ctx->out_verts_cap) return bs_fail(ctx, BS_ERR_INVALID, "no extraction result of that size on the device");
    if (n_floats) BS_CUDA(ctx, cudaMemcpyAsync(d_dst, ctx->d_out_verts, n_floats * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    BS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BS_OK;
}
bs_status bs_ipc_alloc(bs_context* ctx, size_t bytes, void** d_ptr, unsigned char handle[64]) {
    if (!ctx || !d_ptr || !handle) return BS_ERR_INVALID;
    BS_ENTER(ctx);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    *d_ptr = nullptr;
    BS_CUDA(ctx, cudaMalloc(d_ptr, bytes ? bytes : 256));
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, *d_ptr);
    if (e != cudaSuccess) { cudaFree(*d_ptr); *d_ptr = nullptr; return bs_fail(ctx, BS_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
    memcpy(handle, &h, 64);
    return BS_OK;
}
bs_status bs_ipc_open(bs_context* ctx, const unsigned char handle[64], void** d_ptr) {
    if (!ctx || !d_ptr || !handle) return BS_ERR_INVALID;
    BS_ENTER(ctx);
    cudaIpcMemHandle_t h; memcpy(&h, handle, 64);
    BS_CUDA(ctx, cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return BS_OK;
}
bs_status bs_ipc_close(bs_context* ctx, void* d_ptr) {
    if (!ctx) return BS_ERR_INVALID;
    BS_ENTER(ctx);
    if (d_ptr) { BS_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); BS_CUDA(ctx, cudaIpcCloseMemHandle(d_ptr)); }
    return BS_OK;
}
bs_status bs_ipc_free(bs_context* ctx, void* d_ptr) {
    if (!ctx) return BS_ERR_INVALID;
    BS_ENTER(ctx);
    if (d_ptr) { BS_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); BS_CUDA(ctx, cudaFree(d_ptr)); }
    return BS_OK;
}
bs_status bs_context_push_out_verts(bs_context* ctx, float* const* dst, int world, size_t offset_floats, size_t n_floats) {
    if (!ctx || !dst || world < 1 || world > 16) return BS_ERR_INVALID;
    BS_ENTER(ctx);
    if (n_floats > ctx->out_verts_cap) return bs_fail(ctx, BS_ERR_INVALID, "no extraction result of that size on the device");
    PushDst D;
    for (int d = 0; d < 16; ++d) D.p[d] = d < world ? (dst[d] ? dst[d] + offset_floats : nullptr) : nullptr;
    for (int d = 0; d < world; ++d) if (!D.p[d]) return bs_fail(ctx, BS_ERR_INVALID, "null destination %d", d);
    if (n_floats) {
        // all destinations share the offset, and allocations are 256 B aligned: one phase for every destination
        const unsigned mis = (unsigned)(((uintptr_t)D.p[0] >> 2) & 3u), head = mis ? 4u - mis : 0u;
        for (int d = 1; d < world; ++d) if ((((uintptr_t)D.p[d] >> 2) & 3u) != mis) return bs_fail(ctx, BS_ERR_INVALID, "destination buffers must be 16-byte aligned at their base");
        const unsigned grid = (unsigned)std::min<size_t>(bs_blocks(n_floats / 4 + 8, 256), (size_t)ctx->sm_count * 16);
        bs_count_launch(), k_push_out_verts<<<grid, 256, 0, ctx->stream>>>(ctx->d_out_verts, D, world, n_floats, head);
    }
    BS_CUDA(ctx, cudaGetLastError());
    return BS_OK;
}
bs_status bs_context_copy_out_verts(bs_context* ctx, float* dst, size_t n_floats) {
    if (!ctx || (!dst && n_floats)) return BS_ERR_INVALID;
    BS_ENTER(ctx);  // before any bs_fail: a failure belongs to THIS call's epoch
    if (n_floats > ctx->out_verts_cap) return bs_fail(ctx, BS_ERR_INVALID, "no extraction result of that size on the device");
    if (n_floats) BS_CUDA(ctx, cudaMemcpyAsync(dst, ctx->d_out_verts, n_floats * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    BS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BS_OK;
}
size_t bs_context_last_stats(const bs_context* ctx, const char** names, double* values, size_t cap) {
    if (!ctx) return 0;
    size_t n = ctx->stats.size() < cap ? ctx->stats.size() : cap;
    for (size_t i = 0; i < n; ++i) { names[i] = ctx->stats[i].name; values[i] = ctx->stats[i].value; }
    return n;
}

void bs_volume_free(bs_volume* v) {
    if (!v) return;
    bs_context* c = v->ctx;
    if (!c) { delete v; return; }  // orphaned by bs_context_destroy
    BS_ENTER(c);
    c->live_volumes.erase(v);
    if (c->mc_pending.vol == v) bs_mc_pending_release(c);
    bs_free(c, v->keys); bs_free(c, v->values); bs_free(c, v->masks); bs_free(c, v->owned);
    bs_free(c, v->tile8_keys); bs_free(c, v->tile8_values); bs_free(c, v->tile128_keys); bs_free(c, v->tile128_values);
    delete v;
}
float bs_volume_voxel_size(const bs_volume* v) { return v ? v->voxel_size : 0.0f; }

bs_status bs_volume_empty(bs_context* ctx, float voxel_size, bs_volume** out) {
    if (!ctx || !out) return BS_ERR_INVALID;
    BS_ENTER(ctx);
    bs_volume* v = bs_volume_new(ctx, voxel_size);
    bs_status s = bs_volume_alloc_bricks(v, 0);
    if (s != BS_OK) { bs_volume_free(v); return s; }
    *out = v;
    return BS_OK;
}

bs_status bs_volume_clone(const bs_volume* v, bs_volume** out) {
    if (!v || !v->ctx || !out) return BS_ERR_INVALID;
    bs_context* c = v->ctx;
    BS_ENTER(c);
    bs_volume* w = bs_volume_new(c, v->voxel_size);
    w->n_bricks = v->n_bricks; w->n_owned = v->n_owned; w->n_tiles8 = v->n_tiles8; w->n_tiles128 = v->n_tiles128;
    bs_status s;
    if ((v->owned && (s = dup_array(c, &w->owned, v->owned, v->n_bricks))) || (s = dup_array(c, &w->keys, v->keys, v->n_bricks)) || (s = dup_array(c, &w->values, v->values, v->n_bricks * 512)) ||
        (s = dup_array(c, &w->masks, v->masks, v->n_bricks * 8)) || (s = dup_array(c, &w->tile8_keys, v->tile8_keys, v->n_tiles8)) ||
        (s = dup_array(c, &w->tile8_values, v->tile8_values, v->n_tiles8)) || (s = dup_array(c, &w->tile128_keys, v->tile128_keys, v->n_tiles128)) ||
        (s = dup_array(c, &w->tile128_values, v->tile128_values, v->n_tiles128))) { bs_volume_free(w); return s; }
    BS_CUDA(c, cudaStreamSynchronize(c->stream));
    *out = w;
    return BS_OK;
}

void bs_buffer_free(void* p) { if (p) cudaFreeHost(p); }

// ---- data formats either side of the path (bs_io.cu) ---------------------------------------------------------------------------
void bs_device_free(bs_context* ctx, void* d_ptr) { if (ctx && d_ptr) { BS_ENTER(ctx); bs_raw_free(ctx, d_ptr); } }
bs_status bs_stl_decode_device(bs_context* ctx, const unsigned char* d_stl, size_t n_bytes, float** d_tris, size_t* n_tris) {
    if (!ctx || !d_tris || !n_tris || (!d_stl && n_bytes)) return BS_ERR_INVALID;
    BS_ENTER(ctx);
    return bs_stl_decode_impl(ctx, d_stl, n_bytes, d_tris, n_tris);
}
bs_status bs_stl_decode(bs_context* ctx, const unsigned char* stl, size_t n_bytes, float** d_tris, size_t* n_tris) {
    if (!ctx || !d_tris || !n_tris || (!stl && n_bytes)) return BS_ERR_INVALID;
    BS_ENTER(ctx);
    unsigned char* d = nullptr;
    BS_TRY(bs_alloc(ctx, &d, n_bytes + 4));
    if (n_bytes && cudaMemcpyAsync(d, stl, n_bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { bs_free(ctx, d); return bs_fail(ctx, BS_ERR_CUDA, "host to device copy failed"); }
    const bs_status s = bs_stl_decode_impl(ctx, d, n_bytes, d_tris, n_tris);
    bs_free(ctx, d);
    return s;
}
bs_status bs_stl_encode_device(bs_context* ctx, const float* d_verts, size_t n_verts, unsigned char** d_stl, size_t* n_bytes) {
    if (!ctx || !d_stl || !n_bytes || (!d_verts && n_verts)) return BS_ERR_INVALID;
    BS_ENTER(ctx);
    return bs_stl_encode_impl(ctx, d_verts, n_verts, d_stl, n_bytes);
}
bs_status bs_stl_encode(bs_context* ctx, const float* d_verts, size_t n_verts, unsigned char** stl, size_t* n_bytes) {
    if (!ctx || !stl || !n_bytes || (!d_verts && n_verts)) return BS_ERR_INVALID;
    BS_ENTER(ctx);
    unsigned char* d = nullptr;
    BS_TRY(bs_stl_encode_impl(ctx, d_verts, n_verts, &d, n_bytes));
    const bs_status s = to_pinned(ctx, d, *n_bytes, stl);
    bs_free(ctx, d);
    return s;
}
bs_status bs_mesh_active_voxels_device(const bs_volume* v, int32_t** d_verts, size_t* n_verts) {
    if (!v || !v->ctx || !d_verts || !n_verts) return BS_ERR_INVALID;
    BS_ENTER(v->ctx);
    return bs_active_voxels_impl(v, d_verts, n_verts);
}
bs_status bs_mesh_active_voxels(const bs_volume* v, int32_t** verts, size_t* n_verts) {
    if (!v || !v->ctx || !verts || !n_verts) return BS_ERR_INVALID;
    BS_ENTER(v->ctx);
    int* d = nullptr;
    BS_TRY(bs_active_voxels_impl(v, &d, n_verts));
    const bs_status s = to_pinned(v->ctx, d, *n_verts * 3, verts);
    bs_free(v->ctx, d);
    return s;
}
bs_status bs_merge_points_device(bs_context* ctx, const float* d_points, size_t n, float** d_unique, size_t* n_unique, uint32_t** d_indices) {
    if (!ctx || !d_unique || !n_unique || !d_indices || (!d_points && n)) return BS_ERR_INVALID;
    BS_ENTER(ctx);
    return bs_merge_points_impl(ctx, d_points, n, d_unique, n_unique, d_indices);
}
bs_status bs_copy_to_host(bs_context* ctx, const void* d_src, void* dst, size_t bytes) {
    if (!ctx || ((!d_src || !dst) && bytes)) return BS_ERR_INVALID;
    BS_ENTER(ctx);
    if (bytes) BS_CUDA(ctx, cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    BS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BS_OK;
}
bs_status bs_mesh_mc_indexed_device(const bs_volume* v, float voxel_size, float** d_points, size_t* n_points, uint32_t** d_indices, size_t* n_indices) {
    if (!v || !v->ctx || !d_points || !n_points || !d_indices || !n_indices) return BS_ERR_INVALID;
    bs_context* ctx = v->ctx;
    BS_ENTER(ctx);
    const float* d_soup = nullptr; size_t n = 0;
    BS_TRY(bs_mc_impl(v, voxel_size, &d_soup, &n));
    *n_indices = n;
    return bs_merge_points_impl(ctx, d_soup, n, d_points, n_points, d_indices);
}
bs_status bs_mesh_mc_indexed(const bs_volume* v, float voxel_size, float** points, size_t* n_points, uint32_t** indices, size_t* n_indices) {
    if (!v || !v->ctx || !points || !n_points || !indices || !n_indices) return BS_ERR_INVALID;
    bs_context* ctx = v->ctx;
    BS_ENTER(ctx);
    const float* d_soup = nullptr; size_t n = 0;
    BS_TRY(bs_mc_impl(v, voxel_size, &d_soup, &n));
    float* d_u = nullptr; unsigned* d_i = nullptr;
    BS_TRY(bs_merge_points_impl(ctx, d_soup, n, &d_u, n_points, &d_i));
    bs_status s = to_pinned(ctx, d_u, *n_points * 3, points);
    if (s == BS_OK) { s = to_pinned(ctx, d_i, n, indices); if (s != BS_OK) { cudaFreeHost(*points); *points = nullptr; } }
    bs_free(ctx, d_u); bs_free(ctx, d_i);
    *n_indices = n;
    return s;
}
bs_status bs_merge_points(bs_context* ctx, const float* points, size_t n, float** unique, size_t* n_unique, uint32_t** indices) {
    if (!ctx || !unique || !n_unique || !indices || (!points && n)) return BS_ERR_INVALID;
    BS_ENTER(ctx);
    float* d_p = nullptr; float* d_u = nullptr; unsigned* d_i = nullptr;
    BS_TRY(bs_alloc(ctx, &d_p, n * 3));
    if (n && cudaMemcpyAsync(d_p, points, n * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { bs_free(ctx, d_p); return bs_fail(ctx, BS_ERR_CUDA, "host to device copy failed"); }
    bs_status s = bs_merge_points_impl(ctx, d_p, n, &d_u, n_unique, &d_i);
    bs_free(ctx, d_p);
    if (s != BS_OK) return s;
    s = to_pinned(ctx, d_u, *n_unique * 3, unique);
    if (s == BS_OK) { s = to_pinned(ctx, d_i, n, indices); if (s != BS_OK) { cudaFreeHost(*unique); *unique = nullptr; } }
    bs_free(ctx, d_u); bs_free(ctx, d_i);
    return s;
}


bs_status bs_volume_download(const bs_volume* v, int32_t** brick_ijk, float** values, uint64_t** masks, size_t* n_bricks,
                             int32_t** tile_ijk, int32_t** tile_size, float** tile_values, size_t* n_tiles) {
    if (!v || !v->ctx || !brick_ijk || !values || !masks || !n_bricks) return BS_ERR_INVALID;
    bs_context* c = v->ctx;
    BS_ENTER(c);
    const size_t n = v->n_bricks;
    unsigned long long* hkeys = nullptr;
    BS_TRY(to_host(c, &hkeys, v->keys, n));
    BS_TRY(to_host(c, values, v->values, n * 512));
    BS_TRY(to_host(c, (unsigned long long**)masks, v->masks, n * 8));
    unsigned long long *t8k = nullptr, *t128k = nullptr; float *t8v = nullptr, *t128v = nullptr;
    BS_TRY(to_host(c, &t8k, v->tile8_keys, v->n_tiles8)); BS_TRY(to_host(c, &t8v, v->tile8_values, v->n_tiles8));
    BS_TRY(to_host(c, &t128k, v->tile128_keys, v->n_tiles128)); BS_TRY(to_host(c, &t128v, v->tile128_values, v->n_tiles128));
    BS_CUDA(c, cudaStreamSynchronize(c->stream));
    BS_CUDA(c, cudaMallocHost((void**)brick_ijk, (n ? n : 1) * 3 * sizeof(int32_t)));
    for (size_t i = 0; i < n; ++i) {
        int bx, by, bz; bs_key_brick(hkeys[i], bx, by, bz);
        (*brick_ijk)[3 * i] = bx * 8; (*brick_ijk)[3 * i + 1] = by * 8; (*brick_ijk)[3 * i + 2] = bz * 8;
    }
    *n_bricks = n;
    if (tile_ijk && tile_size && tile_values && n_tiles) {
        // merge the two tile lists into the reference's visit order (a 128^3 tile sorts by its node4 key)
        const size_t nt = v->n_tiles8 + v->n_tiles128;
        BS_CUDA(c, cudaMallocHost((void**)tile_ijk, (nt ? nt : 1) * 3 * sizeof(int32_t)));
        BS_CUDA(c, cudaMallocHost((void**)tile_size, (nt ? nt : 1) * sizeof(int32_t)));
        BS_CUDA(c, cudaMallocHost((void**)tile_values, (nt ? nt : 1) * sizeof(float)));
        size_t i8 = 0, i128 = 0, o = 0;
        while (i8 < v->n_tiles8 || i128 < v->n_tiles128) {
            bool take8 = i128 >= v->n_tiles128 || (i8 < v->n_tiles8 && t8k[i8] < (t128k[i128] << 12));
            int bx, by, bz;
            if (take8) { bs_key_brick(t8k[i8], bx, by, bz); (*tile_size)[o] = 8; (*tile_values)[o] = t8v[i8]; ++i8; }
            else { bs_key_brick(t128k[i128] << 12, bx, by, bz); (*tile_size)[o] = 128; (*tile_values)[o] = t128v[i128]; ++i128; }
            (*tile_ijk)[3 * o] = bx * 8; (*tile_ijk)[3 * o + 1] = by * 8; (*tile_ijk)[3 * o + 2] = bz * 8;
            ++o;
        }
        *n_tiles = nt;
    }
    cudaFreeHost(hkeys); cudaFreeHost(t8k); cudaFreeHost(t8v); cudaFreeHost(t128k); cudaFreeHost(t128v);
    return BS_OK;
}

// ---- mesh -> volume ---------------------------------------------------------------------------------------
bs_status bs_mesh_to_volume_sharded(bs_context* ctx, const float* d_tris, size_t n_tris, float voxel_size, int64_t band,
                                    int rank, int world, bs_volume** out) {
    if (!ctx || !out || world < 1 || rank < 0 || rank >= world) return BS_ERR_INVALID;
    *out = nullptr;
    BS_ENTER(ctx);  // before any bs_fail: a failure belongs to THIS call's epoch
    if (!(voxel_size > 0.0f) || band < 0 || band > 64) return bs_fail(ctx, BS_ERR_INVALID, "voxel_size must be > 0 and 0 <= band_width <= 64");
    if (n_tris == 0) return BS_ERR_EMPTY_MESH;
    if (!d_tris) return bs_fail(ctx, BS_ERR_INVALID, "null triangle pointer");
    if (!(voxel_size < 1.0e30f)) return bs_fail(ctx, BS_ERR_RANGE, "voxel_size must be below 1e30");  // (edge vectors of in-range triangles stay finite)
    return bs_convert_impl(ctx, d_tris, n_tris, voxel_size, band, rank, world, out);
}
bs_status bs_mesh_to_volume_device(bs_context* ctx, const float* d_tris, size_t n_tris, float voxel_size, int64_t band, bs_volume** out) {
    return bs_mesh_to_volume_sharded(ctx, d_tris, n_tris, voxel_size, band, 0, 1, out);
}
bs_status bs_mesh_to_volume(bs_context* ctx, const float* tris, size_t n_tris, float voxel_size, int64_t band, bs_volume** out) {
    if (!ctx || !out) return BS_ERR_INVALID;
    *out = nullptr;
    if (n_tris == 0) return BS_ERR_EMPTY_MESH;
    if (!tris) return BS_ERR_INVALID;
    BS_ENTER(ctx);
    float* d = nullptr;
    BS_TRY(bs_alloc(ctx, &d, n_tris * 9));
    cudaError_t e = cudaMemcpyAsync(d, tris, n_tris * 9 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { bs_free(ctx, d); return bs_fail(ctx, BS_ERR_CUDA, "H2D triangles: %s", cudaGetErrorString(e)); }
    bs_status s = bs_mesh_to_volume_device(ctx, d, n_tris, voxel_size, band, out);
    bs_free(ctx, d);
    cudaStreamSynchronize(ctx->stream);
    return s;
}

// ---- extraction ---------------------------------------------------------------------------------------------
static bs_status verts_to_host(bs_context* c, const float* d, size_t n, float** verts, size_t* n_verts) {
    *verts = nullptr; *n_verts = 0;
    BS_CUDA(c, cudaMallocHost((void**)verts, (n ? n : 1) * 3 * sizeof(float)));
    if (n) BS_CUDA(c, cudaMemcpyAsync(*verts, d, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    BS_CUDA(c, cudaStreamSynchronize(c->stream));
    *n_verts = n;
    return BS_OK;
}
bs_status bs_mesh_mc_device(const bs_volume* v, float voxel_size, const float** d_verts, size_t* n_verts) {
    if (!v || !v->ctx || !d_verts || !n_verts) return BS_ERR_INVALID;
    BS_ENTER(v->ctx);
    return bs_mc_impl(v, voxel_size, d_verts, n_verts);
}
bs_status bs_mesh_mc_count(const bs_volume* v, float voxel_size, size_t* n_verts) {
    if (!v || !v->ctx || !n_verts) return BS_ERR_INVALID;
    bs_context* ctx = v->ctx;
    BS_ENTER(ctx);
    if (v->n_tiles8 || v->n_tiles128) return bs_fail(ctx, BS_ERR_UNSUPPORTED, "two-step marching cubes on a volume with active tiles (use bs_mesh_mc_device)");
    bs_mc_pending_release(ctx);
    bs_marks_begin(ctx);
    const bs_status s = bs_mc_count_phase(v, voxel_size, n_verts);
    if (s != BS_OK) return s;
    bs_marks_end(ctx);
    bs_stat_add(ctx, "n_bricks", (double)v->n_bricks);
    bs_stat_add(ctx, "n_out_tris", (double)(*n_verts / 3));
    return BS_OK;
}
bs_status bs_mesh_mc_emit_push(const bs_volume* v, float* const* dst, int world, size_t offset_floats, size_t cap_floats) {
    if (!v || !v->ctx || !dst || world < 1 || world > 16) return BS_ERR_INVALID;
    bs_context* ctx = v->ctx;
    BS_ENTER(ctx);
    for (int d = 0; d < world; ++d) if (!dst[d]) return bs_fail(ctx, BS_ERR_INVALID, "null destination %d", d);
    bs_marks_begin(ctx);
    const bs_status s = bs_mc_emit_phase(v, dst, world, offset_floats, cap_floats);
    bs_mc_pending_release(ctx);
    if (s != BS_OK) return s;
    BS_CUDA(ctx, cudaGetLastError());
    bs_marks_end(ctx);
    return BS_OK;
}
bs_status bs_mesh_mc(const bs_volume* v, float voxel_size, float** verts, size_t* n_verts) {
    if (!v || !v->ctx || !verts || !n_verts) return BS_ERR_INVALID;
    BS_ENTER(v->ctx);  // extraction + read-back under one lock
    const float* d = nullptr; size_t n = 0;
    BS_TRY(bs_mesh_mc_device(v, voxel_size, &d, &n));
    return verts_to_host(v->ctx, d, n, verts, n_verts);
}
bs_status bs_mesh_dc_device(const bs_volume* v, float voxel_size, const float** d_verts, size_t* n_verts) {
    if (!v || !v->ctx || !d_verts || !n_verts) return BS_ERR_INVALID;
    BS_ENTER(v->ctx);
    return bs_dc_impl(v, voxel_size, d_verts, n_verts);
}
bs_status bs_mesh_dc(const bs_volume* v, float voxel_size, float** verts, size_t* n_verts) {
    if (!v || !v->ctx || !verts || !n_verts) return BS_ERR_INVALID;
    BS_ENTER(v->ctx);  // extraction + read-back under one lock
    const float* d = nullptr; size_t n = 0;
    BS_TRY(bs_mesh_dc_device(v, voxel_size, &d, &n));
    return verts_to_host(v->ctx, d, n, verts, n_verts);
}

// ---- VoxelRemesher::remesh, host triangles in, host vertices out -----------------------------------------------
// The mesh is uploaded once and converted + extracted slab by slab (the brick slabs of the multi-GPU path, run one after
// the other on this GPU): the device-to-host copy of slab k's vertices runs on a second stream while slab k + 1 is being
// converted, so the 1.1 GB read-back of a 10 M-triangle remesh hides behind the kernels instead of following them. Slabs are
// contiguous ranges of the reference's leaf visit order, so the concatenation is the single-pass result bit for bit; the
// one piece of state the reference carries from cell to cell (the MC33 c-vertex) is handed from slab
// to slab on the device (d_mc_carry).
static void remesh_accumulate(std::vector<bs_stat>& acc, const std::vector<bs_stat>& st) {
    for (const bs_stat& a : st) {
        const size_t len = strlen(a.name);
        const bool additive = (len > 3 && !strcmp(a.name + len - 3, "_ms")) || !strcmp(a.name, "n_active") || !strcmp(a.name, "n_out_tris") || !strcmp(a.name, "n_eval") || !strcmp(a.name, "n_bricks_owned") || !strcmp(a.name, "n_sub") || !strcmp(a.name, "n_sign_seeds");
        bool found = false;
        for (bs_stat& b : acc) if (!strcmp(a.name, b.name)) { if (additive) b.value += a.value; else b.value = a.value; found = true; break; }
        if (!found) acc.push_back(a);
    }
}
bs_status bs_voxel_remesh_into(bs_context* ctx, const float* tris, size_t n_tris, float voxel_size, int method, int slabs,
                               float* dst, size_t cap_floats, size_t* n_floats) {
    if (!ctx || !n_floats || (!dst && cap_floats)) return BS_ERR_INVALID;
    *n_floats = 0;
    if (n_tris == 0) return BS_ERR_EMPTY_MESH;
    if (!tris || (method != 0 && method != 1)) return BS_ERR_INVALID;
    BS_ENTER(ctx);
    if (!(voxel_size > 0.0f) || !(voxel_size < 1.0e30f)) return bs_fail(ctx, BS_ERR_INVALID, "voxel_size must be in (0, 1e30)");
    int K = slabs;
    if (K <= 0) { if (const char* e = getenv("BSHARK_REMESH_SLABS")) K = atoi(e); }
    if (K <= 0) K = n_tris >= (1u << 20) ? 4 : 1;  // below ~1 M triangles the per-slab overhead outweighs the hidden copy
    if (K > 64) K = 64;
    cudaStream_t st = ctx->stream;
    if (!ctx->copy_stream) {
        BS_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) { BS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming)); BS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming)); }
        BS_CUDA(ctx, cudaMalloc((void**)&ctx->d_mc_carry, 4 * sizeof(float)));
    }
    float* d = nullptr;
    BS_TRY(bs_alloc(ctx, &d, n_tris * 9));
    BS_CUDA(ctx, cudaMemcpyAsync(d, tris, n_tris * 9 * sizeof(float), cudaMemcpyHostToDevice, st));
    BS_CUDA(ctx, cudaMemsetAsync(ctx->d_mc_carry, 0, 4 * sizeof(float), st));
    ctx->mc_chain = K > 1 && method == 0;
    bs_convert_plan plan;
    std::vector<bs_stat> acc;
    size_t off = 0; int n_copies = 0; bs_status s = BS_OK; bool any_bricks = false;
    const size_t CHUNK = (size_t)8 << 20;  // floats per copy: small control read-backs of the running slab slip in between
    for (int k = 0; k < K && s == BS_OK; ++k) {
        bs_volume* v = nullptr;
        s = bs_convert_impl(ctx, d, n_tris, voxel_size, 0, k, K, &v, &plan);
        if (s != BS_OK) break;
        any_bricks = any_bricks || v->n_bricks != 0;
        remesh_accumulate(acc, ctx->stats);
        // the buffer this extraction writes was the source of copy k - 2
        if (n_copies >= 2) cudaStreamWaitEvent(st, ctx->ev_copied[k & 1], 0);
        const float* dv = nullptr; size_t nv = 0;
        s = method == 0 ? bs_mc_impl(v, voxel_size, &dv, &nv) : bs_dc_impl(v, voxel_size, &dv, &nv);
        bs_volume_free(v);
        if (s != BS_OK) break;
        remesh_accumulate(acc, ctx->stats);
        const size_t nf = nv * 3;
        if (nf && off + nf <= cap_floats) {
            cudaEventRecord(ctx->ev_done[k & 1], st);
            cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[k & 1], 0);
            for (size_t c = 0; c < nf; c += CHUNK) {
                const cudaError_t e = cudaMemcpyAsync(dst + off + c, dv + c, std::min(CHUNK, nf - c) * sizeof(float), cudaMemcpyDeviceToHost, ctx->copy_stream);
                if (e != cudaSuccess) { s = bs_fail(ctx, BS_ERR_CUDA, "D2H vertices: %s", cudaGetErrorString(e)); break; }
            }
        }
        cudaEventRecord(ctx->ev_copied[k & 1], ctx->copy_stream);
        ++n_copies;
        off += nf;
        std::swap(ctx->d_out_verts, ctx->d_out_alt); std::swap(ctx->out_verts_cap, ctx->out_alt_cap);  // the next slab extracts into the other buffer
    }
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamSynchronize(st);
    ctx->mc_chain = false;
    bs_free(ctx, d);
    if (s != BS_OK) return s;
    if (!any_bricks) return BS_ERR_EMPTY_MESH;  // nothing of the mesh produced a voxel in any slab: convert -> None (mesh_to_volume.rs:58-60)
    ctx->stats = acc;
    bs_stat_add(ctx, "remesh_slabs", (double)K);
    *n_floats = off;
    if (off > cap_floats) return bs_fail(ctx, BS_ERR_INVALID, "result buffer too small: %zu floats needed, %zu given (n_floats holds the size to retry with)", off, cap_floats);
    return BS_OK;
}

// ---- CSG / offset -------------------------------------------------------------------------------------------
static bs_status csg_entry(bs_volume* a, bs_volume* b, int op, bs_volume** out) {
    if (out) *out = nullptr;
    if (!a || !b || !out || a == b || !a->ctx || a->ctx != b->ctx) { bs_volume_free(a); if (b != a) bs_volume_free(b); return BS_ERR_INVALID; }
    BS_ENTER(a->ctx);
    if (a->owned || b->owned) { bs_status e = bs_fail(a->ctx, BS_ERR_UNSUPPORTED, "CSG on a brick-sharded volume (DESIGN.md, Multi-GPU)"); bs_volume_free(a); bs_volume_free(b); return e; }
    bs_status s = bs_csg_impl(a, b, op, out);
    bs_volume_free(a); bs_volume_free(b);
    return s;
}
bs_status bs_volume_union(bs_volume* a, bs_volume* b, bs_volume** out) { return csg_entry(a, b, 0, out); }
bs_status bs_volume_subtract(bs_volume* a, bs_volume* b, bs_volume** out) { return csg_entry(a, b, 1, out); }
bs_status bs_volume_intersect(bs_volume* a, bs_volume* b, bs_volume** out) { return csg_entry(a, b, 2, out); }
bs_status bs_volume_offset(bs_volume* a, float distance, bs_volume** out) {
    if (out) *out = nullptr;
    if (!a || !out || !a->ctx) { bs_volume_free(a); return BS_ERR_INVALID; }
    BS_ENTER(a->ctx);
    if (a->owned) { bs_status e = bs_fail(a->ctx, BS_ERR_UNSUPPORTED, "offset does not shard: the sweep wavefronts cross slabs (DESIGN.md, Multi-GPU)"); bs_volume_free(a); return e; }
    bs_status s = bs_offset_impl(a, distance, out);
    bs_volume_free(a);
    return s;
}

// ---- builders -----------------------------------------------------------------------------------------------
bs_status bs_volume_from_voxels(bs_context* ctx, const int32_t* ijk, const float* values, size_t m, float voxel_size, bs_volume** out) {
    if (!ctx || !out || (m && (!ijk || !values))) return BS_ERR_INVALID;
    *out = nullptr;
    BS_ENTER(ctx);
    if (m == 0) return bs_volume_empty(ctx, voxel_size, out);
    int32_t* d_ijk = nullptr; float* d_val = nullptr;
    BS_TRY(bs_alloc(ctx, &d_ijk, m * 3)); BS_TRY(bs_alloc(ctx, &d_val, m));
    cudaMemcpyAsync(d_ijk, ijk, m * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d_val, values, m * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
    bs_status s = bs_from_voxels_impl(ctx, d_ijk, d_val, m, voxel_size, out);
    bs_free(ctx, d_ijk); bs_free(ctx, d_val);
    cudaStreamSynchronize(ctx->stream);
    return s;
}
bs_status bs_volume_sphere(bs_context* ctx, float voxel_size, float radius, const float origin[3], bs_volume** out) {
    if (!ctx || !out || !origin) return BS_ERR_INVALID;
    BS_ENTER(ctx);
    float p[7] = {radius, origin[0], origin[1], origin[2], 0, 0, 0};
    return bs_builder_impl(ctx, 0, voxel_size, p, out);
}
bs_status bs_volume_cuboid(bs_context* ctx, float voxel_size, const float mn[3], const float mx[3], bs_volume** out) {
    if (!ctx || !out || !mn || !mx) return BS_ERR_INVALID;
    BS_ENTER(ctx);
    float p[7] = {mn[0], mn[1], mn[2], mx[0], mx[1], mx[2], 0};
    return bs_builder_impl(ctx, 1, voxel_size, p, out);
}
bs_status bs_volume_iwp(bs_context* ctx, float voxel_size, const float mn[3], const float mx[3], float cell_size, bs_volume** out) {
    if (!ctx || !out || !mn || !mx) return BS_ERR_INVALID;
    BS_ENTER(ctx);
    float p[7] = {mn[0], mn[1], mn[2], mx[0], mx[1], mx[2], cell_size};
    return bs_builder_impl(ctx, 2, voxel_size, p, out);
}

}  // extern "C"
