// baby_shark.hpp -- header-only C++17 host mirror of baby_shark's `voxel::prelude` and `remeshing::voxel` over the
// C ABI in bshark.h. Same names, argument meaning and error behaviour as the Rust API it mirrors:
//   None            -> std::nullopt           (MeshToVolume::convert on an empty mesh, mesh_to_volume.rs:58-60)
//   panic           -> baby_shark::Panic      (todo!() / unwrap() / unreachable!() sites, see bshark.h)
//   move semantics  -> rvalue-qualified union_/subtract/intersect/offset consume *this and the argument
// (the reference is compiled Rust; no Rust toolchain exists in this image, so the runnable host mirror is C++.)
#pragma once
#include <algorithm>
#include <array>
#include <cstddef>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "bshark.h"

namespace baby_shark {

using Vec3f = std::array<float, 3>;  // helpers/aliases.rs:3-7

struct Error : std::runtime_error { bs_status status; Error(bs_status s, const std::string& m) : std::runtime_error(m), status(s) {} };
struct Panic : Error { using Error::Error; };  // the reference panics on this input

class Context {
public:
    explicit Context(int device = -1) { bs_status s = bs_context_create(device, &h_); if (s != BS_OK) throw Error(s, "bs_context_create: a B200 is required (no CPU fallback)"); }
    ~Context() { bs_context_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    bs_context* get() const { return h_; }
    void check(bs_status s) const {
        if (s == BS_OK) return;
        if (s == BS_ERR_REFERENCE_PANICS) throw Panic(s, bs_last_error(h_));
        throw Error(s, bs_last_error(h_));
    }
    static Context& global() { static Context c; return c; }
private:
    bs_context* h_ = nullptr;
};

namespace voxel {

// voxel::volume::Volume (src/voxel/volume/mod.rs:10-123)
class Volume {
public:
    Volume(Volume&& o) noexcept : h_(std::exchange(o.h_, nullptr)), ctx_(o.ctx_) {}
    Volume& operator=(Volume&& o) noexcept { if (this != &o) { bs_volume_free(h_); h_ = std::exchange(o.h_, nullptr); ctx_ = o.ctx_; } return *this; }
    ~Volume() { bs_volume_free(h_); }
    static Volume with_voxel_size(float voxel_size, Context& c = Context::global()) { bs_volume* h; c.check(bs_volume_empty(c.get(), voxel_size, &h)); return Volume(h, &c); }
    // from_fn (volume/mod.rs:40-72): the closure is evaluated on the host exactly as the reference does
    template <class F>
    static Volume from_fn(float voxel_size, Vec3f min, Vec3f max, std::size_t narrow_band_width, F func, Context& c = Context::global()) {
        const float nbw = float(narrow_band_width + 1) * voxel_size;
        long lo[3], hi[3];
        for (int d = 0; d < 3; ++d) { lo[d] = (long)__builtin_floorf(min[d] / voxel_size); hi[d] = (long)__builtin_ceilf(max[d] / voxel_size); }
        std::vector<int32_t> ijk; std::vector<float> val;
        for (long x = lo[0]; x <= hi[0]; ++x) for (long y = lo[1]; y <= hi[1]; ++y) for (long z = lo[2]; z <= hi[2]; ++z) {
            const Vec3f p{float(x) * voxel_size, float(y) * voxel_size, float(z) * voxel_size};
            const float v = func(p);
            if (__builtin_fabsf(v) > nbw) continue;
            ijk.insert(ijk.end(), {int32_t(x), int32_t(y), int32_t(z)}); val.push_back(v);
        }
        bs_volume* h; c.check(bs_volume_from_voxels(c.get(), ijk.data(), val.data(), val.size(), voxel_size, &h));
        return Volume(h, &c);
    }
    float voxel_size() const { return bs_volume_voxel_size(h_); }
    Volume clone() const { bs_volume* h; ctx_->check(bs_volume_clone(h_, &h)); return Volume(h, ctx_); }
    // consuming, like `fn union(mut self, mut other: Self) -> Self` (volume/mod.rs:74-93); `union` is a C++ keyword
    Volume union_(Volume other) && { return binary(bs_volume_union, std::move(other)); }
    Volume subtract(Volume other) && { return binary(bs_volume_subtract, std::move(other)); }
    Volume intersect(Volume other) && { return binary(bs_volume_intersect, std::move(other)); }
    Volume offset(float distance) && {  // volume/mod.rs:95-108
        bs_volume* h; bs_volume* a = std::exchange(h_, nullptr);
        ctx_->check(bs_volume_offset(a, distance, &h));
        return Volume(h, ctx_);
    }
    bs_volume* raw() const { return h_; }
    Context& context() const { return *ctx_; }
    Volume(bs_volume* h, Context* c) : h_(h), ctx_(c) {}
private:
    template <class Fn> Volume binary(Fn fn, Volume other) {
        bs_volume* h; bs_volume *a = std::exchange(h_, nullptr), *b = std::exchange(other.h_, nullptr);
        ctx_->check(fn(a, b, &h));
        return Volume(h, ctx_);
    }
    bs_volume* h_ = nullptr;
    Context* ctx_ = nullptr;
};

// voxel::mesh_to_volume::MeshToVolume (src/voxel/mesh_to_volume.rs:17-73; Default :223-236)
class MeshToVolume {
public:
    explicit MeshToVolume(Context& c = Context::global()) : ctx_(&c) {}
    MeshToVolume& with_narrow_band_width(long width) { band_width_ = width; return *this; }
    MeshToVolume& set_narrow_band_width(long width) { band_width_ = width; return *this; }
    MeshToVolume& with_voxel_size(float size) { voxel_size_ = size; return *this; }
    MeshToVolume& set_voxel_size(float size) { voxel_size_ = size; return *this; }
    // `mesh`: n x 9 floats, the triangles `Triangles::triangles()` yields (mesh/traits.rs:4-8)
    std::optional<Volume> convert(const float* tris, std::size_t n_tris) {
        bs_volume* h = nullptr;
        const bs_status s = bs_mesh_to_volume(ctx_->get(), tris, n_tris, voxel_size_, band_width_, &h);
        if (s == BS_ERR_EMPTY_MESH) return std::nullopt;
        ctx_->check(s);
        return Volume(h, ctx_);
    }
    std::optional<Volume> convert(const std::vector<Vec3f>& soup) { return convert(soup.empty() ? nullptr : soup[0].data(), soup.size() / 3); }
private:
    Context* ctx_; float voxel_size_ = 1.0f; long band_width_ = 0;
};

// voxel::volume::builder::VolumeBuilder (src/voxel/volume/builder.rs:5-84)
class VolumeBuilder {
public:
    explicit VolumeBuilder(Context& c = Context::global()) : ctx_(&c) {}
    VolumeBuilder& with_voxel_size(float v) { voxel_size_ = v; return *this; }
    void set_voxel_size(float v) { voxel_size_ = v; }
    Volume sphere(float radius, Vec3f origin) const { bs_volume* h; ctx_->check(bs_volume_sphere(ctx_->get(), voxel_size_, radius, origin.data(), &h)); return Volume(h, ctx_); }
    Volume cuboid(Vec3f min, Vec3f max) const { bs_volume* h; ctx_->check(bs_volume_cuboid(ctx_->get(), voxel_size_, min.data(), max.data(), &h)); return Volume(h, ctx_); }
    Volume iwp(Vec3f min, Vec3f max, float cell_size) const { bs_volume* h; ctx_->check(bs_volume_iwp(ctx_->get(), voxel_size_, min.data(), max.data(), cell_size, &h)); return Volume(h, ctx_); }
private:
    Context* ctx_; float voxel_size_ = 1.0f;
};

namespace detail {
inline std::vector<Vec3f> take(Context& c, const float* d_verts, std::size_t n_verts) {
    std::vector<Vec3f> out(n_verts);  // fill the caller-visible Vec directly: no library-owned staging buffer
    (void)d_verts;
    c.check(bs_context_copy_out_verts(c.get(), n_verts ? out[0].data() : nullptr, n_verts * 3));
    return out;
}
}  // namespace detail

// voxel::meshing::MarchingCubesMesher (src/voxel/meshing/marching_cubes.rs:17-63)
class MarchingCubesMesher {
public:
    MarchingCubesMesher& with_voxel_size(float s) { voxel_size_ = s; return *this; }
    MarchingCubesMesher& set_voxel_size(float s) { voxel_size_ = s; return *this; }
    std::vector<Vec3f> mesh(const Volume& sdf) {
        const float* d; std::size_t n;
        sdf.context().check(bs_mesh_mc_device(sdf.raw(), voxel_size_, &d, &n));
        return detail::take(sdf.context(), d, n);
    }
private:
    float voxel_size_ = 1.0f;
};

// voxel::meshing::DualContouringMesher (src/voxel/meshing/dual_contouring.rs:13-89)
class DualContouringMesher {
public:
    DualContouringMesher& with_voxel_size(float s) { voxel_size_ = s; return *this; }
    std::optional<std::vector<Vec3f>> mesh(const Volume& volume) {
        const float* d; std::size_t n;
        volume.context().check(bs_mesh_dc_device(volume.raw(), voxel_size_, &d, &n));  // throws Panic where the reference panics
        return detail::take(volume.context(), d, n);
    }
private:
    float voxel_size_ = 1.0f;
};

}  // namespace voxel

// io::stl (src/io/stl.rs): binary STL decoded / encoded on the device
namespace io {
// n x 9 f32 triangles resident on the device (library-owned); feeds MeshToVolume::convert without a host round trip
class DeviceTriangles {
public:
    DeviceTriangles(float* d, std::size_t n, Context* c) : d_(d), n_(n), ctx_(c) {}
    DeviceTriangles(DeviceTriangles&& o) noexcept : d_(std::exchange(o.d_, nullptr)), n_(o.n_), ctx_(o.ctx_) {}
    DeviceTriangles(const DeviceTriangles&) = delete;
    ~DeviceTriangles() { if (d_) bs_device_free(ctx_->get(), d_); }
    const float* data() const { return d_; }
    std::size_t size() const { return n_; }
    Context& context() const { return *ctx_; }
private:
    float* d_; std::size_t n_; Context* ctx_;
};
class StlReader {  // src/io/stl.rs:14-95
public:
    explicit StlReader(Context& c = Context::global()) : ctx_(&c) {}
    // throws Error where the reference returns Err(ReadError) (buffer shorter than the header announces)
    DeviceTriangles read_from_buffer(const unsigned char* bytes, std::size_t n_bytes) {
        float* d = nullptr; std::size_t n = 0;
        ctx_->check(bs_stl_decode(ctx_->get(), bytes, n_bytes, &d, &n));
        return DeviceTriangles(d, n, ctx_);
    }
private:
    Context* ctx_;
};
class StlWriter {  // src/io/stl.rs:110-191, for the vertex soup of the last *_device extraction or any device soup
public:
    explicit StlWriter(Context& c = Context::global()) : ctx_(&c) {}
    std::vector<unsigned char> write_to_buffer(const float* d_verts, std::size_t n_verts) {
        unsigned char* h = nullptr; std::size_t n = 0;
        ctx_->check(bs_stl_encode(ctx_->get(), d_verts, n_verts, &h, &n));
        std::vector<unsigned char> out(h, h + n);
        bs_buffer_free(h);
        return out;
    }
private:
    Context* ctx_;
};
}  // namespace io

namespace voxel {
inline std::optional<Volume> convert(MeshToVolume& m2v, const io::DeviceTriangles& t, float voxel_size, long band = 0) {
    (void)m2v;
    bs_volume* h = nullptr;
    const bs_status s = bs_mesh_to_volume_device(t.context().get(), t.data(), t.size(), voxel_size, band, &h);
    if (s == BS_ERR_EMPTY_MESH) return std::nullopt;
    t.context().check(s);
    return Volume(h, &t.context());
}
// voxel::meshing::ActiveVoxelsMesher (src/voxel/meshing/active_voxels.rs:4-22); the reference returns Vector3<isize>
class ActiveVoxelsMesher {
public:
    std::vector<std::array<long, 3>> mesh(const Volume& volume) {
        int32_t* h = nullptr; std::size_t n = 0;
        volume.context().check(bs_mesh_active_voxels(volume.raw(), &h, &n));
        std::vector<std::array<long, 3>> out(n);
        for (std::size_t i = 0; i < n; ++i) out[i] = {h[3 * i], h[3 * i + 1], h[3 * i + 2]};
        bs_buffer_free(h);
        return out;
    }
};
}  // namespace voxel

namespace algo {
// algo::merge_points (src/algo/merge_points.rs:4-41)
struct IndexedVertices { std::vector<Vec3f> points; std::vector<std::size_t> indices; };
inline IndexedVertices merge_points(const std::vector<Vec3f>& points, Context& c = Context::global()) {
    float* u = nullptr; uint32_t* idx = nullptr; std::size_t nu = 0;
    c.check(bs_merge_points(c.get(), points.empty() ? nullptr : points[0].data(), points.size(), &u, &nu, &idx));
    IndexedVertices out;
    out.points.resize(nu);
    for (std::size_t i = 0; i < nu; ++i) out.points[i] = {u[3 * i], u[3 * i + 1], u[3 * i + 2]};
    out.indices.assign(idx, idx + points.size());
    bs_buffer_free(u); bs_buffer_free(idx);
    return out;
}
}  // namespace algo

namespace remeshing {
enum class MeshingMethod { FeaturePreserving, Manifold };  // remeshing/voxel.rs:10-15

// remeshing::voxel::VoxelRemesher (src/remeshing/voxel.rs:45-95)
class VoxelRemesher {
public:
    VoxelRemesher& with_voxel_size(float size) { m2v_.set_voxel_size(size); voxel_size_ = size; return *this; }
    VoxelRemesher& with_meshing_method(MeshingMethod m) { method_ = m; return *this; }
    // one call, host triangles in, host vertices out: upload, slab-wise convert + extraction, read-back overlapped with
    // the kernels (bs_voxel_remesh_into); the same vertices, bit for bit, as convert() followed by mesh()
    std::optional<std::vector<Vec3f>> remesh(const float* tris, std::size_t n_tris, Context& c = Context::global()) {
        std::vector<Vec3f> out(std::max<std::size_t>(1024, 12 * n_tris));
        for (int attempt = 0; attempt < 2; ++attempt) {
            std::size_t n = 0;
            const bs_status st = bs_voxel_remesh_into(c.get(), tris, n_tris, voxel_size_, method_ == MeshingMethod::FeaturePreserving ? 1 : 0, 0, out[0].data(), out.size() * 3, &n);
            if (st == BS_ERR_EMPTY_MESH) return std::nullopt;
            if (st == BS_ERR_INVALID && n > out.size() * 3) { out.resize(n / 3); continue; }  // the estimate was too small: retry with the reported size
            c.check(st);
            out.resize(n / 3);
            return out;
        }
        c.check(BS_ERR_INVALID);
        return std::nullopt;
    }
private:
    voxel::MeshToVolume m2v_{}; MeshingMethod method_ = MeshingMethod::Manifold; float voxel_size_ = 1.0f;
};
}  // namespace remeshing

}  // namespace baby_shark
