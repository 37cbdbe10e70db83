// ORACLE -- TEST INFRASTRUCTURE ONLY. C ABI over the CPU restatement in bso_*.h so that tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs can drive it through
// ctypes. The product (baby_shark_b200/) never links or calls this library.
//
// Parity pinning (see DESIGN.md "Oracle"): pinned against the reference's own known answers
//   - volume/mod.rs:134-152 (7944 MC vertices for box2.stl @0.2 offset 0.5),
//   - leaf_node/csg.rs:53-132, leaf_node/flood_fill.rs:82-113, internal_node/flood_fill.rs:128-193,
//     root_node/flood_fill.rs:75-95, voxel/tests.rs:10-95 (bso_selftest below),
// everything else (winding numbers, tree build, MC33 ambiguous branches, DC, internal/root CSG,
// builders) is "parity unpinned by the reference": the reference has no tests for it and cannot be
// built here (no Rust toolchain), so this restatement is the only pin.
#include "bso_io.h"
#include <chrono>
#include <stdexcept>
#include <cstdio>
#include <cstdlib>

namespace bso {
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}
using namespace bso;

extern "C" {

void* bso_mesh_to_volume(const float* tris, size_t n, float voxel_size, int64_t band, int threads, ConvertStats* st, int count_work) {
    VolumeGrid* g = mesh_to_volume(tris, n, voxel_size, band, threads, st, count_work != 0);
    if (!g) return nullptr;
    return new Volume{g, voxel_size};
}
void* bso_volume_sphere(float vs, float r, float ox, float oy, float oz) { return volume_sphere(vs, r, Vec3f{ox, oy, oz}); }
void* bso_volume_cuboid(float vs, const float* mn, const float* mx) { return volume_cuboid(vs, Vec3f{mn[0], mn[1], mn[2]}, Vec3f{mx[0], mx[1], mx[2]}); }
void* bso_volume_iwp(float vs, const float* mn, const float* mx, float cell) { return volume_iwp(vs, Vec3f{mn[0], mn[1], mn[2]}, Vec3f{mx[0], mx[1], mx[2]}, cell); }
void* bso_volume_empty(float vs) { return new Volume{new VolumeGrid(), vs}; }
void* bso_volume_from_voxels(const int32_t* ijk, const float* val, size_t m, float vs) {
    VolumeGrid* g = new VolumeGrid();
    for (size_t i = 0; i < m; ++i) g->insert(Vec3i{ijk[3 * i], ijk[3 * i + 1], ijk[3 * i + 2]}, val[i]);
    return new Volume{g, vs};
}
void* bso_volume_clone(void* v) { Volume* a = (Volume*)v; return new Volume{a->grid->clone(), a->voxel_size}; }
void bso_volume_free(void* v) { if (!v) return; Volume* a = (Volume*)v; delete a->grid; delete a; }
float bso_volume_voxel_size(void* v) { return ((Volume*)v)->voxel_size; }

// volume/mod.rs:74-93 ; `a` is updated in place, `b` is consumed
static void csg(void* a_, void* b_, int op) {
    Volume *a = (Volume*)a_, *b = (Volume*)b_;
    a->grid->flood_fill(); b->grid->flood_fill();
    if (op == 0) a->grid->csg_union(b->grid); else if (op == 1) a->grid->csg_subtract(b->grid); else a->grid->csg_intersect(b->grid);
    delete b;
}
void bso_volume_union(void* a, void* b) { csg(a, b, 0); }
void bso_volume_subtract(void* a, void* b) { csg(a, b, 1); }
void bso_volume_intersect(void* a, void* b) { csg(a, b, 2); }
void bso_volume_flood_fill(void* a) { ((Volume*)a)->grid->flood_fill(); }
int bso_volume_sign_at(void* a, int64_t x, int64_t y, int64_t z) { return ((Volume*)a)->grid->sign_at(Vec3i{x, y, z}) == Negative ? -1 : 1; }

int bso_volume_offset(void* v, float distance, SweepStats* st) {
    try { volume_offset((Volume*)v, distance, st); } catch (std::exception& e) { return 1; }
    return 0;
}

struct CountVisitor { size_t leaves = 0, active = 0, tiles = 0, negative = 0;
    void dense(const Leaf3<float>& l) { leaves++; for (int i = 0; i < 512; ++i) if (l.value_mask.at(i)) { active++; if (std::signbit(l.values[i])) negative++; } }
    void tile(const Tile<float>&) { tiles++; } };
void bso_volume_counts(void* v, size_t* n_leaves, size_t* n_active, size_t* n_tiles, size_t* n_negative) {
    CountVisitor c; ((Volume*)v)->grid->visit_leafs(c);
    if (n_leaves) *n_leaves = c.leaves;
    if (n_active) *n_active = c.active;
    if (n_tiles) *n_tiles = c.tiles;
    if (n_negative) *n_negative = c.negative;
}
struct DownloadVisitor { int32_t* origins; float* values; uint64_t* masks; int32_t* tile_origins; int32_t* tile_sizes; float* tile_values; size_t nl = 0, nt = 0;
    void dense(const Leaf3<float>& l) {
        Vec3i o = l.origin(); origins[3 * nl] = int32_t(o.x); origins[3 * nl + 1] = int32_t(o.y); origins[3 * nl + 2] = int32_t(o.z);
        std::memcpy(values + 512 * nl, l.values, 512 * sizeof(float)); std::memcpy(masks + 8 * nl, l.value_mask.w, 64); nl++; }
    void tile(const Tile<float>& t) { tile_origins[3 * nt] = int32_t(t.origin.x); tile_origins[3 * nt + 1] = int32_t(t.origin.y); tile_origins[3 * nt + 2] = int32_t(t.origin.z); tile_sizes[nt] = int32_t(t.size); tile_values[nt] = t.value; nt++; } };
// leaves and tiles each in the reference's visit order (root map order, ascending slot offsets)
void bso_volume_download(void* v, int32_t* origins, float* values, uint64_t* masks, int32_t* tile_origins, int32_t* tile_sizes, float* tile_values) {
    DownloadVisitor d{origins, values, masks, tile_origins, tile_sizes, tile_values};
    ((Volume*)v)->grid->visit_leafs(d);
}

static float* to_buffer(const std::vector<Vec3f>& v, size_t* n) {
    *n = v.size();
    float* out = (float*)std::malloc(std::max<size_t>(1, v.size()) * 3 * sizeof(float));
    for (size_t i = 0; i < v.size(); ++i) { out[3 * i] = v[i].x; out[3 * i + 1] = v[i].y; out[3 * i + 2] = v[i].z; }
    return out;
}
int bso_mesh_mc(void* v, float voxel_size, float** verts, size_t* n_verts, McStats* st) {
    MarchingCubes mc; mc.mesh(*(Volume*)v, voxel_size);
    *verts = to_buffer(mc.vertices, n_verts);
    if (st) *st = mc.stats;
    return 0;
}
int bso_mesh_dc(void* v, float voxel_size, float** verts, size_t* n_verts) {
    DualContouring dc; const char* err = nullptr;
    if (!dc.mesh(*(Volume*)v, voxel_size, &err)) { *verts = nullptr; *n_verts = 0; return 1; }
    *verts = to_buffer(dc.out, n_verts);
    return 0;
}
void bso_buffer_free(float* p) { std::free(p); }

// stage-level hooks for kernel parity tests
size_t bso_subdivide(const float* tris, size_t n, float voxel_size, float** out) {
    std::vector<Tri> sub;
    for (size_t i = 0; i < n; ++i) { const float* p = tris + 9 * i; subdivide_triangle(Tri{{p[0], p[1], p[2]}, {p[3], p[4], p[5]}, {p[6], p[7], p[8]}}, voxel_size, sub); }
    *out = (float*)std::malloc(std::max<size_t>(1, sub.size()) * 9 * sizeof(float));
    std::memcpy(*out, sub.data(), sub.size() * 9 * sizeof(float));
    return sub.size();
}
void bso_point_triangle_distance(const float* tri9, const float* pts, size_t m, float* out) {
    Tri t{{tri9[0], tri9[1], tri9[2]}, {tri9[3], tri9[4], tri9[5]}, {tri9[6], tri9[7], tri9[8]}};
    for (size_t i = 0; i < m; ++i) { Vec3f p{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]}; out[i] = norm(closest_point(t, p) - p); }
}
// approximate (beta > 0) or exact (beta <= 0) winding numbers on the reference's SAH tree
void bso_winding_numbers(const float* tris, size_t n, const float* pts, size_t m, float beta, float* out, uint64_t* counters) {
    WindingNumbers wn; wn.build(tris, n);
    WindingNumbers::Counters cnt;
    for (size_t i = 0; i < m; ++i) {
        Vec3f p{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
        if (beta > 0.0f) out[i] = wn.approximate(p, beta, counters ? &cnt : nullptr);
        else { float w = 0.0f; for (auto& o : wn.objects) w += WindingNumbers::solid_angle(o.first, p); out[i] = w / (4.0f * 3.14159265358979323846f); }
    }
    if (counters) { counters[0] = cnt.n_visit; counters[1] = cnt.n_far; counters[2] = cnt.n_exact; counters[3] = wn.nodes.size(); }
}
// data formats either side of the path
int bso_stl_decode(const uint8_t* bytes, size_t n_bytes, float** tris, size_t* n_tris) {
    std::vector<float> t;
    if (!stl_decode(bytes, n_bytes, t)) { *tris = nullptr; *n_tris = 0; return 1; }
    *n_tris = t.size() / 9;
    *tris = (float*)std::malloc(std::max<size_t>(1, t.size()) * sizeof(float));
    std::memcpy(*tris, t.data(), t.size() * sizeof(float));
    return 0;
}
void bso_stl_encode(const float* verts, size_t n_tris, uint8_t* out /* 84 + 50 n_tris bytes */) {
    std::vector<uint8_t> o; stl_encode(verts, n_tris, o);
    std::memcpy(out, o.data(), o.size());
}
size_t bso_active_voxels(void* v, int32_t** out) {
    ActiveVoxels av{((Volume*)v)->grid, {}};
    ((Volume*)v)->grid->visit_leafs(av);
    *out = (int32_t*)std::malloc(std::max<size_t>(1, av.out.size()) * 3 * sizeof(int32_t));
    for (size_t i = 0; i < av.out.size(); ++i) { (*out)[3 * i] = (int32_t)av.out[i].x; (*out)[3 * i + 1] = (int32_t)av.out[i].y; (*out)[3 * i + 2] = (int32_t)av.out[i].z; }
    return av.out.size();
}
size_t bso_merge_points(const float* pts, size_t n, float* unique /* up to 3 n */, uint32_t* indices /* n */) {
    std::vector<float> u; std::vector<uint32_t> idx;
    merge_points(pts, n, u, idx);
    std::memcpy(unique, u.data(), u.size() * sizeof(float));
    std::memcpy(indices, idx.data(), idx.size() * sizeof(uint32_t));
    return u.size() / 3;
}
void bso_free(void* p) { std::free(p); }
float bso_compute_distance(float a1, float a2, float a3, float h) { return compute_distance(a1, a2, a3, h); }

// ------------------------------------------------------------------------------------------------
// The reference's own unit tests for this path, restated on the oracle's node classes.
#define CHECK(cond) do { if (!(cond)) { std::fprintf(stderr, "bso_selftest failed: %s (line %d)\n", #cond, __LINE__); return __LINE__; } } while (0)
int bso_selftest() {
    {   // leaf_node/csg.rs:53-132
        typedef LeafNode<float, 1> L;
        const float v1[8] = {10, 20, 30, 40, -10, -20, -30, -40}, v2[8] = {5, 15, 25, 35, -5, -15, -25, -35};
        const float un[8] = {5, 15, 25, 35, -10, -20, -30, -40}, su[8] = {10, 20, 30, 40, 5, 15, 25, 35}, in[8] = {10, 20, 30, 40, -5, -15, -25, -35};
        for (int op = 0; op < 3; ++op) {
            L *a = L::empty({0, 0, 0}), *b = L::empty({0, 0, 0});
            std::memcpy(a->values, v1, sizeof(v1)); std::memcpy(b->values, v2, sizeof(v2)); a->value_mask.on_all(); b->value_mask.on_all();
            if (op == 0) a->csg_union(b); else if (op == 1) a->csg_subtract(b); else a->csg_intersect(b);
            const float* exp = op == 0 ? un : op == 1 ? su : in;
            for (int i = 0; i < 8; ++i) CHECK(a->values[i] == exp[i]);
            CHECK(a->value_mask.is_full());
            delete a;
        }
    }
    {   // leaf_node/flood_fill.rs:82-113
        typedef LeafNode<float, 2> L;
        L* n = L::empty({0, 0, 0}); n->insert({2, 2, 2}, 1.0f); n->flood_fill();
        for (int i = 0; i < 64; ++i) CHECK(!std::signbit(n->values[i]));
        delete n;
        n = L::empty({0, 0, 0}); n->insert({2, 2, 2}, -1.0f); n->flood_fill();
        for (int i = 0; i < 64; ++i) CHECK(std::signbit(n->values[i]));
        delete n;
        n = L::empty({0, 0, 0});
        for (idx_t y = 0; y < 4; ++y) for (idx_t z = 0; z < 4; ++z) { n->insert({1, y, z}, -1.0f); n->insert({2, y, z}, 1.0f); }
        n->flood_fill();
        for (int i = 0; i < 32; ++i) CHECK(std::signbit(n->values[i]));
        for (int i = 32; i < 64; ++i) CHECK(!std::signbit(n->values[i]));
        delete n;
    }
    {   // internal_node/flood_fill.rs:128-193 : static_vdb!(f32, 2, 1)
        typedef LeafNode<float, 1> L; typedef InternalNode<float, L, 2> I;
        const idx_t R = idx_t(I::resolution());
        CHECK(R == 8);
        I* n = I::empty({0, 0, 0}); n->insert({3, 2, 1}, 1.0f); n->flood_fill();
        for (idx_t x = 0; x < R; ++x) for (idx_t y = 0; y < R; ++y) for (idx_t z = 0; z < R; ++z) CHECK(n->sign_at({x, y, z}) == Positive);
        n->destroy();
        n = I::empty({0, 0, 0}); n->insert({3, 3, 3}, -1.0f); n->flood_fill();
        for (idx_t x = 0; x < R; ++x) for (idx_t y = 0; y < R; ++y) for (idx_t z = 0; z < R; ++z) CHECK(n->sign_at({x, y, z}) == Negative);
        n->destroy();
        n = I::empty({0, 0, 0});
        for (idx_t y = 0; y < R; ++y) for (idx_t z = 0; z < R; ++z) { n->insert({4, y, z}, -1.0f); n->insert({5, y, z}, 1.0f); }
        n->flood_fill();
        for (idx_t x = 0; x < 5; ++x) for (idx_t y = 0; y < 5; ++y) for (idx_t z = 0; z < 5; ++z) CHECK(n->sign_at({x, y, z}) == Negative);
        for (idx_t x = 5; x < R; ++x) for (idx_t y = 5; y < R; ++y) for (idx_t z = 5; z < R; ++z) CHECK(n->sign_at({x, y, z}) == Positive);
        n->destroy();
        n = I::empty({0, 0, 0});
        for (idx_t x = 2; x < 4; ++x) for (idx_t y = 0; y < R; ++y) for (idx_t z = 0; z < R; ++z) n->insert({x, y, z}, -1.0f);
        for (idx_t x = 4; x < 6; ++x) for (idx_t y = 0; y < R; ++y) for (idx_t z = 0; z < R; ++z) n->insert({x, y, z}, 1.0f);
        n->flood_fill();
        for (idx_t x = 0; x < 4; ++x) for (idx_t y = 0; y < 4; ++y) for (idx_t z = 0; z < 4; ++z) CHECK(n->sign_at({x, y, z}) == Negative);
        for (idx_t x = 4; x < R; ++x) for (idx_t y = 4; y < R; ++y) for (idx_t z = 4; z < R; ++z) CHECK(n->sign_at({x, y, z}) == Positive);
        n->destroy();
    }
    {   // root_node/flood_fill.rs:75-95 : dynamic_vdb!(f32, 1)
        typedef RootNode<LeafNode<float, 1>> T;
        T t;
        for (idx_t x = 0; x < 2; ++x) for (idx_t y = 4; y < 6; ++y) for (idx_t z = 4; z < 6; ++z) t.insert({x, y, z}, -1.0f);
        for (idx_t x = 0; x < 2; ++x) for (idx_t y = 4; y < 6; ++y) for (idx_t z = 10; z < 12; ++z) t.insert({x, y, z}, -1.0f);
        t.flood_fill();
        for (idx_t x = 0; x < 2; ++x) for (idx_t y = 4; y < 6; ++y) for (idx_t z = 4; z < 12; ++z) CHECK(t.sign_at({x, y, z}) == Negative);
    }
    {   // voxel/tests.rs:10-95 (Empty value type -> char): static and dynamic 4,3,2 trees
        typedef LeafNode<char, 2> L; typedef InternalNode<char, L, 3> I3; typedef InternalNode<char, I3, 4> I4;
        I4* t = I4::empty({0, 0, 0});
        for (idx_t x = 0; x < 32; ++x) for (idx_t y = 0; y < 32; ++y) for (idx_t z = 0; z < 32; ++z) { t->insert({x, y, z}, 0); CHECK(t->at({x, y, z})); }
        CHECK(!t->is_empty());
        t->remove_if([](char) { return true; });
        CHECK(t->is_empty());
        t->destroy();
        RootNode<I4> d;
        CHECK(d.is_empty());
        for (idx_t x = 0; x < 32; ++x) for (idx_t y = 0; y < 32; ++y) for (idx_t z = 0; z < 32; ++z) { d.insert({x, y, z}, 0); CHECK(d.at({x, y, z})); }
        CHECK(!d.is_empty());
        typedef InternalNode<char, LeafNode<char, 2>, 3> F;  // static_vdb!(Empty, 3, 2) fill test
        F* f = F::empty({0, 0, 0}); CHECK(f->is_empty()); f->fill(0);
        for (idx_t x = 0; x < 32; ++x) for (idx_t y = 0; y < 32; ++y) for (idx_t z = 0; z < 32; ++z) CHECK(f->at({x, y, z}));
        f->destroy();
    }
    return 0;
}

}  // extern "C"
