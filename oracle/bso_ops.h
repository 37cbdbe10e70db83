// ORACLE -- TEST INFRASTRUCTURE ONLY (see bso_tree.h). CPU restatement of
//   src/voxel/volume/mod.rs:40-108 (from_fn, union/intersect/subtract, offset),
//   src/voxel/volume/builder.rs:21-76 (sphere, cuboid, iwp),
//   src/voxel/fast_sweep.rs:30-287,509-537 (block Gauss-Seidel fast sweeping),
//   src/helpers/utils.rs:4-16 (sort3), src/voxel/utils.rs:63-73 (option_min_by).
#pragma once
#include "bso_convert.h"
#include <set>
#include <stdexcept>

namespace bso {

struct Volume { VolumeGrid* grid; float voxel_size; };

// volume/mod.rs:40-72
template <class F> Volume* volume_from_fn(float voxel_size, Vec3f mn, Vec3f mx, size_t narrow_band_width, F func) {
    VolumeGrid* grid = new VolumeGrid();
    float nbw = float(narrow_band_width + 1) * voxel_size;
    Vec3i lo{f2i(std::floor(mn.x / voxel_size)), f2i(std::floor(mn.y / voxel_size)), f2i(std::floor(mn.z / voxel_size))};
    Vec3i hi{f2i(std::ceil(mx.x / voxel_size)), f2i(std::ceil(mx.y / voxel_size)), f2i(std::ceil(mx.z / voxel_size))};
    for (idx_t x = lo.x; x <= hi.x; ++x) for (idx_t y = lo.y; y <= hi.y; ++y) for (idx_t z = lo.z; z <= hi.z; ++z) {
        Vec3f p{float(x) * voxel_size, float(y) * voxel_size, float(z) * voxel_size};
        float v = func(p);
        if (std::fabs(v) > nbw) continue;  // NaN is kept, like the reference
        grid->insert(Vec3i{x, y, z}, v);
    }
    return new Volume{grid, voxel_size};
}
inline Vec3f add_scalar(Vec3f v, float s) { return {v.x + s, v.y + s, v.z + s}; }

inline Volume* volume_sphere(float vs, float radius, Vec3f origin) {  // builder.rs:21-30
    size_t bw = 1;
    float off = radius + float(bw) * vs;
    return volume_from_fn(vs, add_scalar(origin, -off), add_scalar(origin, off), bw, [=](const Vec3f& p) { return norm(p - origin) - radius; });
}
inline Volume* volume_cuboid(float vs, Vec3f mn, Vec3f mx) {  // builder.rs:32-51
    size_t bw = 1;
    float off = float(bw) * vs;
    Box3f box{mn, mx};
    return volume_from_fn(vs, add_scalar(mn, -off), add_scalar(mx, off), bw, [=](const Vec3f& p) {
        if (box.contains_point(p)) {
            float m = std::fmin(p.x - mn.x, mx.x - p.x);
            m = std::fmin(m, p.y - mn.y); m = std::fmin(m, mx.y - p.y);
            m = std::fmin(m, p.z - mn.z); m = std::fmin(m, mx.z - p.z);
            return -m;
        }
        return std::sqrt(box.squared_distance(p));
    });
}
inline Volume* volume_iwp(float vs, Vec3f mn, Vec3f mx, float cell_size) {  // builder.rs:54-76
    float inv = 1.0f / cell_size;
    Box3f sampling{mn, mx};
    Box3f iwp{add_scalar(mn, -vs), add_scalar(mx, vs)};
    return volume_from_fn(vs, iwp.mn, iwp.mx, 2, [=](const Vec3f& p) {
        float x = p.x * inv, y = p.y * inv, z = p.z * inv;
        float cx = std::cos(x), cy = std::cos(y), cz = std::cos(z);
        float v = -(cx + cy + cz - 0.51f * (cx * cy + cy * cz + cz * cx) - 1.0f);
        if (!sampling.contains_point(p)) return std::sqrt(iwp.squared_distance(p));
        return v * cell_size;
    });
}

// ------------------------------------------------------------------------------------------------
// fast_sweep.rs
inline void sort3(float& a, float& b, float& c) { if (a > c) std::swap(a, c); if (a > b) std::swap(a, b); if (b > c) std::swap(b, c); }
inline float compute_distance(float a1, float a2, float a3, float h) {  // :509-537
    sort3(a1, a2, a3);
    float s1 = a1 + h;
    if (std::fabs(s1) <= a2) return s1;
    float a12 = a1 + a2, hsq = h * h, two = hsq + hsq, d12 = a1 - a2, d12s = d12 * d12;
    float s2 = (a12 + std::sqrt(two - d12s)) * 0.5f;
    if (std::fabs(s2) <= a3) return s2;
    float a123 = a12 + a3, three = two + hsq, d13 = a1 - a3, d13s = d13 * d13, d23 = a2 - a3, d23s = d23 * d23;
    return (a123 + std::sqrt(three - d12s - d13s - d23s)) * (1.0f / 3.0f);
}

struct SweepStats { uint64_t leaves_processed[8]; uint64_t n_leaves_final; };

struct FastSweeping {
    typedef Leaf3<float> Leaf;
    float limit_abs, grid_spacing; Sign sweep_sign;
    Grid<char>* frozen = nullptr;
    SweepStats stats;
    FastSweeping(float spacing, float limit) : limit_abs(std::fabs(limit)), grid_spacing(spacing), sweep_sign(sign_of(limit)) { std::memset(&stats, 0, sizeof(stats)); }
    ~FastSweeping() { delete frozen; }

    struct Stencil {
        Leaf *top, *bottom, *left, *right, *front, *back, *center; const Leaf3<char>* frozen; Vec3i mn, mx;
        const float* at(const Vec3i& i) const {
            if (i.z < mn.z) return bottom->at(i);
            if (i.z >= mx.z) return top->at(i);
            if (i.y < mn.y) return front->at(i);
            if (i.y >= mx.y) return back->at(i);
            if (i.x < mn.x) return left->at(i);
            if (i.x >= mx.x) return right->at(i);
            return center->at(i);
        }
    };
    // option_min_by with cmp_abs: ties (and incomparable) -> first argument
    static const float* min_abs(const float* a, const float* b) {
        if (a && b) { return (std::fabs(*a) > std::fabs(*b)) ? b : a; }
        return a ? a : b;
    }
    void sweep_voxel(const Vec3i& idx, Stencil& st) {
        if (st.frozen && st.frozen->at(idx)) return;
        const float* dx = min_abs(st.at({idx.x + 1, idx.y, idx.z}), st.at({idx.x - 1, idx.y, idx.z}));
        const float* dy = min_abs(st.at({idx.x, idx.y + 1, idx.z}), st.at({idx.x, idx.y - 1, idx.z}));
        const float* dz = min_abs(st.at({idx.x, idx.y, idx.z + 1}), st.at({idx.x, idx.y, idx.z - 1}));
        const float* first = dx ? dx : (dy ? dy : dz);
        if (!first) return;
        if (sign_of(*first) != sweep_sign) return;
        float d1 = dx ? *dx : far_value(), d2 = dy ? *dy : far_value(), d3 = dz ? *dz : far_value();
        float d_new_abs = compute_distance(std::fabs(d1), std::fabs(d2), std::fabs(d3), grid_spacing);
        float d_new = d_new_abs; set_sign(d_new, sweep_sign);
        if (d_new_abs > limit_abs) return;
        const float* old = st.center->at(idx);
        float d_old = old ? *old : far_value();
        if (d_new_abs < std::fabs(d_old)) st.center->insert(idx, d_new);
    }
    // dir bits: bit0 = -x, bit1 = -y, bit2 = -z ; sweep order :39-60 is dir = 0..7
    void sweep(VolumeGrid* sdf, std::vector<Vec3i>& work, int dir) {
        const idx_t size = 8;
        const int sx = (dir & 1) ? -1 : 1, sy = (dir & 2) ? -1 : 1, sz = (dir & 4) ? -1 : 1;
        auto less = [=](const Vec3i& a, const Vec3i& b) {  // processing order: lexicographic (x,y,z) in sweep direction
            if (a.x != b.x) return sx > 0 ? a.x < b.x : a.x > b.x;
            if (a.y != b.y) return sy > 0 ? a.y < b.y : a.y > b.y;
            if (a.z != b.z) return sz > 0 ? a.z < b.z : a.z > b.z;
            return false;
        };
        std::set<Vec3i, decltype(less)> heap(less);      // BinaryHeap popping the smallest
        std::set<Vec3i, decltype(less)> existing(less);  // HashSet of everything ever queued in this sweep
        for (auto& o : work) { heap.insert(o); existing.insert(o); }
        std::vector<Vec3i> removed;
        while (!heap.empty()) {
            Vec3i o = *heap.begin(); heap.erase(heap.begin());
            removed.push_back(o);
            stats.leaves_processed[dir]++;
            auto take = [&](Vec3i q) { Leaf* l = sdf->take_leaf_at(q); return l ? l : Leaf::empty(q); };
            Stencil st;
            st.left = take({o.x - size, o.y, o.z}); st.right = take({o.x + size, o.y, o.z});
            st.front = take({o.x, o.y - size, o.z}); st.back = take({o.x, o.y + size, o.z});
            st.top = take({o.x, o.y, o.z + size}); st.bottom = take({o.x, o.y, o.z - size});
            st.center = sdf->take_leaf_at(o);
            if (!st.center) { throw std::runtime_error("fast_sweep: center leaf missing (reference unwrap() panics; tiles in offset input)"); }
            st.frozen = frozen->leaf_at(o);
            st.mn = o; st.mx = Vec3i{o.x + size, o.y + size, o.z + size};
            // sweep_stencil :111-123
            for (idx_t xi = 0; xi < size; ++xi) for (idx_t yi = 0; yi < size; ++yi) for (idx_t zi = 0; zi < size; ++zi) {
                Vec3i idx{sx > 0 ? o.x + xi : o.x + size - 1 - xi, sy > 0 ? o.y + yi : o.y + size - 1 - yi, sz > 0 ? o.z + zi : o.z + size - 1 - zi};
                sweep_voxel(idx, st);
            }
            // insert_neighboring_nodes :185-287 ; note the negative-direction face index `origin + origin`
            // collapses to local 0 through the leaf's index masking.
            auto face_has = [&](int axis) {
                bool pos = axis == 0 ? sx > 0 : (axis == 1 ? sy > 0 : sz > 0);
                idx_t oc = axis == 0 ? o.x : (axis == 1 ? o.y : o.z);
                idx_t fc = pos ? size - 1 : oc;  // added to origin below, as in the reference
                for (idx_t u = 0; u < size; ++u) for (idx_t v = 0; v < size; ++v) {
                    Vec3i idx;
                    if (axis == 2) idx = Vec3i{u + o.x, v + o.y, fc + o.z};
                    else if (axis == 1) idx = Vec3i{u + o.x, fc + o.y, v + o.z};
                    else idx = Vec3i{fc + o.x, u + o.y, v + o.z};
                    const float* val = st.center->at(idx);
                    if (val && sign_of(*val) == sweep_sign && std::fabs(*val) < limit_abs) return true;
                }
                return false;
            };
            bool iz = face_has(2), iy = face_has(1), ix = face_has(0);
            auto push = [&](Vec3i q) { if (existing.insert(q).second) heap.insert(q); };
            if (ix) push({sx > 0 ? o.x + size : o.x - size, o.y, o.z});
            if (iy) push({o.x, sy > 0 ? o.y + size : o.y - size, o.z});
            if (iz) push({o.x, o.y, sz > 0 ? o.z + size : o.z - size});
            sdf->insert_leaf_at(st.top); sdf->insert_leaf_at(st.bottom); sdf->insert_leaf_at(st.left); sdf->insert_leaf_at(st.right);
            sdf->insert_leaf_at(st.front); sdf->insert_leaf_at(st.back); sdf->insert_leaf_at(st.center);
        }
        work = removed;
    }
    void fast_sweep(VolumeGrid* sdf) {
        delete frozen;
        frozen = sdf->clone_map<char>([](float) { return char(0); });
        struct Collect { std::vector<Vec3i> o; void dense(const Leaf& l) { o.push_back(l.origin()); } void tile(const Tile<float>& t) { o.push_back(t.origin); } } col;
        sdf->visit_leafs(col);
        std::vector<Vec3i> work = col.o;
        for (int dir = 0; dir < 8; ++dir) sweep(sdf, work, dir);
        sdf->remove_empty_branches();
    }
};

inline void volume_offset(Volume* v, float distance, SweepStats* st) {  // volume/mod.rs:95-108
    const float vs = v->voxel_size;
    v->grid->remove_if([vs](float val) { return std::fabs(val) > vs * 2.0f; });
    float ext = std::fabs(distance) + vs + vs;
    set_sign(ext, sign_of(distance));
    FastSweeping sw(vs, ext);
    sw.fast_sweep(v->grid);
    auto sub = [distance](float& val) { val -= distance; };
    v->grid->visit_values_mut(sub);
    if (st) { *st = sw.stats; }
}

}  // namespace bso
