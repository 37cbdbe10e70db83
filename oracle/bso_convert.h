// ORACLE -- TEST INFRASTRUCTURE ONLY (see bso_tree.h). CPU restatement of
//   src/voxel/mesh_to_volume.rs (subdivide :75-116, unsigned field :118-196, signs :198-281),
//   src/geometry/primitives/triangle3.rs:113-124,261-280,307-382 (center, max_side, normal, area,
//   bbox, closest_point), src/geometry/primitives/box3.rs (center/area/offset/union),
//   src/spatial_partitioning/aabb_tree.rs:67-80,118-271,422-514 (top-down SAH build) and
//   :571-817 (fast winding numbers: solid angle, dipole order-1/2 coefficients, traversal).
// nalgebra 0.34 / nalgebra-glm 0.20 are not vendored with the reference; their 3-vector kernels are
// restated as: dot = (x*x' + y*y') + z*z'; norm_squared = (x*x + y*y) + z*z; norm = sqrt(norm_squared);
// normalize = v / norm; cross = (ay*bz - az*by, az*bx - ax*bz, ax*by - ay*bx). Compile with
// -ffp-contract=off so no FMA is formed (Rust never contracts).
#pragma once
#include "bso_tree.h"
#include <thread>
#include <atomic>
#include <mutex>

namespace bso {

inline Vec3f operator+(Vec3f a, Vec3f b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3f operator-(Vec3f a, Vec3f b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3f operator*(Vec3f a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3f operator/(Vec3f a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float dot(Vec3f a, Vec3f b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float norm_squared(Vec3f a) { return (a.x * a.x + a.y * a.y) + a.z * a.z; }
inline float norm(Vec3f a) { return std::sqrt(norm_squared(a)); }
inline Vec3f cross(Vec3f a, Vec3f b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline Vec3f min2(Vec3f a, Vec3f b) { return {std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)}; }
inline Vec3f max2(Vec3f a, Vec3f b) { return {std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)}; }
inline float comp(const Vec3f& v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

struct Tri { Vec3f a, b, c; };
struct Box3f {
    Vec3f mn, mx;
    static Box3f empty() { return {{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}}; }  // box3.rs:22-27 (min_value of f32 = -MAX)
    Vec3f center() const { return (mn + mx) / 2.0f; }
    void union_box(const Box3f& o) { mx = max2(mx, o.mx); mn = min2(mn, o.mn); }
    void union_point(const Vec3f& p) { mx = max2(mx, p); mn = min2(mn, p); }
    float area() const { Vec3f d = mx - mn; return 2.0f * (d.x * d.y + d.x * d.z + d.y * d.z); }
    Vec3f offset(const Vec3f& p) const {  // box3.rs:185-201
        Vec3f o = p - mn;
        if (mx.x > mn.x) o.x /= mx.x - mn.x;
        if (mx.y > mn.y) o.y /= mx.y - mn.y;
        if (mx.z > mn.z) o.z /= mx.z - mn.z;
        return o;
    }
    bool contains_point(const Vec3f& p) const { return p.x >= mn.x && p.x <= mx.x && p.y >= mn.y && p.y <= mx.y && p.z >= mn.z && p.z <= mx.z; }
    float squared_distance(const Vec3f& p) const {  // box3.rs:95-111
        float sq = 0.0f;
        for (int i = 0; i < 3; ++i) {
            float v = comp(p, i), lo = comp(mn, i), hi = comp(mx, i);
            if (v < lo) sq += (lo - v) * (lo - v);
            if (v > hi) sq += (v - hi) * (v - hi);
        }
        return sq;
    }
};
inline Box3f tri_bbox(const Tri& t) { return {min2(t.c, min2(t.a, t.b)), max2(t.c, max2(t.a, t.b))}; }  // triangle3.rs:307-315
inline float tri_max_side(const Tri& t) {  // :117-124
    float ab = norm_squared(t.b - t.a), ac = norm_squared(t.c - t.a), bc = norm_squared(t.c - t.b);
    return std::sqrt(std::fmax(std::fmax(ab, ac), bc));
}
inline Vec3f tri_center(const Tri& t) { return (t.a + t.b + t.c) / 3.0f; }
inline bool tri_is_degenerate(Vec3f a, Vec3f b, Vec3f c) { return norm_squared(cross(b - a, c - a)) == 0.0f; }

// triangle3.rs:317-382 (Ericson)
inline Vec3f closest_point(const Tri& t, const Vec3f& p) {
    Vec3f ab = t.b - t.a, ac = t.c - t.a, ap = p - t.a;
    float d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0.0f && d2 <= 0.0f) return t.a;
    Vec3f bp = p - t.b;
    float d3 = dot(ab, bp), d4 = dot(ac, bp);
    if (d3 >= 0.0f && d4 <= d3) return t.b;
    float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) { float v = d1 / (d1 - d3); return t.a + ab * v; }
    Vec3f cp = p - t.c;
    float d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0.0f && d5 <= d6) return t.c;
    float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) { float w = d2 / (d2 - d6); return t.a + ac * w; }
    float va = d3 * d6 - d5 * d4;
    if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) { float w = (d4 - d3) / ((d4 - d3) + (d5 - d6)); return t.b + (t.c - t.b) * w; }
    float denom = 1.0f / (va + vb + vc);
    float v = vb * denom, w = vc * denom;
    return t.a + ab * v + ac * w;
}

// ------------------------------------------------------------------------------------------------
// AABB tree (aabb_tree.rs:67-271, Area strategy :422-514) + winding numbers (:571-817)
struct WindingNumbers {
    struct Node { bool leaf; size_t left, right; Box3f bbox; };
    struct NodeData { Vec3f o1; float o2[9]; float radius; Vec3f center; };  // o2 column-major m[col*3+row]
    struct Init { Vec3f awn, awc; float o1sum[9]; float area; Vec3f center; };
    std::vector<Node> nodes;
    std::vector<std::pair<Tri, Box3f>> objects;
    std::vector<NodeData> data;
    size_t min_objects_per_leaf = 3, max_depth = 40;
    // traversal counters: only the COUNT instantiation of fast_wn touches them, and then through a
    // caller-owned (thread-local) block -- the timed path has no shared writes
    struct Counters { uint64_t n_visit = 0, n_far = 0, n_exact = 0; };

    void build(const float* tris, size_t n) {
        objects.clear(); nodes.clear(); data.clear();
        objects.reserve(n);
        for (size_t i = 0; i < n; ++i) {
            const float* p = tris + 9 * i;
            Tri t{{p[0], p[1], p[2]}, {p[3], p[4], p[5]}, {p[6], p[7], p[8]}};
            objects.push_back({t, tri_bbox(t)});
        }
        if (!objects.empty()) build_node(0, objects.size(), 1);
        if (!nodes.empty()) { data.resize(nodes.size()); compute_node_data(nodes.size() - 1); }
    }
    size_t leaf_from(size_t first, size_t last) {
        Box3f bb = objects[first].second;
        for (size_t i = first + 1; i < last; ++i) bb.union_box(objects[i].second);
        nodes.push_back({true, first, last, bb});
        return nodes.size() - 1;
    }
    size_t build_node(size_t first, size_t last, size_t depth) {
        if (depth >= max_depth || last - first <= min_objects_per_leaf) return leaf_from(first, last);
        long split = split_objects(first, last);
        if (split < 0) return leaf_from(first, last);
        size_t at = size_t(split) + first;
        size_t l = build_node(first, at, depth + 1);
        size_t r = build_node(at, last, depth + 1);
        Box3f bb = nodes[l].bbox; bb.union_box(nodes[r].bbox);
        nodes.push_back({false, l, r, bb});
        return nodes.size() - 1;
    }
    long split_objects(size_t first, size_t last) {
        Box3f bb = objects[first].second;
        for (size_t i = first; i < last; ++i) bb.union_box(objects[i].second);
        std::pair<float, int> ax[3] = {{bb.mx.x - bb.mn.x, 0}, {bb.mx.y - bb.mn.y, 1}, {bb.mx.z - bb.mn.z, 2}};
        std::stable_sort(ax, ax + 3, [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first > b.first; });
        for (int k = 0; k < 3; ++k) {
            int axis = ax[k].second;
            std::stable_sort(objects.begin() + first, objects.begin() + last,
                             [axis](const std::pair<Tri, Box3f>& a, const std::pair<Tri, Box3f>& b) { return comp(a.second.center(), axis) < comp(b.second.center(), axis); });
            long s = area_split(first, last, axis, bb);
            if (s >= 0) return s;
        }
        return -1;
    }
    static size_t bucket_of(const Box3f& cb, const Vec3f& c, int axis) {
        float f = 12.0f * comp(cb.offset(c), axis);
        // Rust `as usize`: saturating, NaN -> 0
        size_t t = (f != f || f <= 0.0f) ? 0 : (f >= 1.8e19f ? SIZE_MAX : size_t(f));
        return std::min<size_t>(t, 11);
    }
    long area_split(size_t first, size_t last, int axis, const Box3f& objects_bbox) {
        if (first == last) return -1;
        Box3f cb = Box3f::empty();
        for (size_t i = first; i < last; ++i) cb.union_point(objects[i].second.center());
        struct Bucket { size_t count; Box3f bb; };
        Bucket buckets[12];
        for (auto& b : buckets) { b.count = 0; b.bb = Box3f::empty(); }
        for (size_t i = first; i < last; ++i) {
            size_t bi = bucket_of(cb, objects[i].second.center(), axis);
            buckets[bi].count++; buckets[bi].bb.union_box(objects[i].second);
        }
        float costs[11];
        for (int i = 0; i < 11; ++i) {
            Box3f b0 = Box3f::empty(), b1 = Box3f::empty();
            size_t c0 = 0, c1 = 0;
            for (int j = 0; j <= i; ++j) { b0.union_box(buckets[j].bb); c0 += buckets[j].count; }
            for (int j = i + 1; j < 12; ++j) { b1.union_box(buckets[j].bb); c1 += buckets[j].count; }
            float f0 = float(c0), f1 = float(c1);
            if (f0 == 0.0f || f1 == 0.0f) costs[i] = INFINITY;
            else costs[i] = 0.125f + (f0 * b0.area() + f1 * b1.area()) / objects_bbox.area();
        }
        int best = 0;  // Iterator::min_by returns the FIRST minimum
        for (int i = 1; i < 11; ++i) if (costs[i] < costs[best]) best = i;
        float leaf_cost = float(last - first);
        if (costs[best] < leaf_cost) {
            for (size_t i = first; i < last; ++i)
                if (bucket_of(cb, objects[i].second.center(), axis) > size_t(best)) return long(i - first);
            return -1;
        }
        return -1;
    }
    Init compute_node_data(size_t idx) {
        const Node node = nodes[idx];
        Init d;
        if (node.leaf) {
            d.awn = {0, 0, 0}; d.awc = {0, 0, 0}; d.area = 0.0f;
            for (int i = 0; i < 9; ++i) d.o1sum[i] = 0.0f;
            for (size_t t = node.left; t < node.right; ++t) {
                const Tri& tri = objects[t].first;
                Vec3f cr = cross(tri.b - tri.a, tri.c - tri.a);
                if (norm_squared(cr) == 0.0f) continue;
                Vec3f n = cr / norm(cr);
                float area = norm(cr) * 0.5f;
                d.area += area;
                d.awn = d.awn + n * area;
                Vec3f c = tri_center(tri);
                Vec3f ac = c * area;
                float cv[3] = {ac.x, ac.y, ac.z}, nv[3] = {n.x, n.y, n.z};
                for (int col = 0; col < 3; ++col) for (int row = 0; row < 3; ++row) d.o1sum[col * 3 + row] += cv[row] * nv[col];
                d.awc = d.awc + ac;
            }
            d.center = d.awc / d.area;
        } else {
            Init l = compute_node_data(node.left), r = compute_node_data(node.right);
            for (int i = 0; i < 9; ++i) d.o1sum[i] = l.o1sum[i] + r.o1sum[i];
            d.awn = l.awn + r.awn; d.awc = l.awc + r.awc; d.area = l.area + r.area;
            d.center = (l.awc + r.awc) / d.area;
        }
        float dmin = norm_squared(node.bbox.mn - d.center), dmax = norm_squared(node.bbox.mx - d.center);
        NodeData nd;
        nd.radius = std::sqrt(std::fmax(dmin, dmax));
        nd.o1 = d.awn; nd.center = d.center;
        float cv[3] = {d.center.x, d.center.y, d.center.z}, nv[3] = {d.awn.x, d.awn.y, d.awn.z};
        for (int col = 0; col < 3; ++col) for (int row = 0; row < 3; ++row) nd.o2[col * 3 + row] = d.o1sum[col * 3 + row] - cv[row] * nv[col];
        data[idx] = nd;
        return d;
    }
    static float solid_angle(const Tri& t, const Vec3f& q) {  // :582-615
        Vec3f qa = t.a - q, qb = t.b - q, qc = t.c - q;
        float al = norm(qa), bl = norm(qb), cl = norm(qc);
        if (al == 0.0f || bl == 0.0f || cl == 0.0f) return 0.0f;
        qa = qa / al; qb = qb / bl; qc = qc / cl;
        float num = dot(qa, cross(qb - qa, qc - qa));
        if (num == 0.0f) return 0.0f;
        float den = 1.0f + dot(qa, qb) + dot(qa, qc) + dot(qb, qc);
        return std::atan2(num, den) * 2.0f;
    }
    float approximate(const Vec3f& p, float beta, Counters* c = nullptr) const {
        if (nodes.empty()) return 0.0f;
        return c ? fast_wn<true>(nodes.size() - 1, p, beta, c) : fast_wn<false>(nodes.size() - 1, p, beta, nullptr);
    }
    template <bool COUNT> float fast_wn(size_t idx, const Vec3f& p, float beta, Counters* c) const {
        const NodeData& nd = data[idx];
        if (COUNT) c->n_visit++;
        float dist = norm(p - nd.center);
        if (dist > nd.radius * beta) {
            if (COUNT) c->n_far++;
            const float PI = 3.14159265358979323846f;
            Vec3f r = nd.center - p;
            float r2 = norm_squared(r), r1 = std::sqrt(r2), r3 = r2 * r1;
            float den = 4.0f * PI * r3, inv = 1.0f / den;
            Vec3f ord1 = r * inv;
            float r5 = r3 * r2;
            float rv[3] = {r.x, r.y, r.z};
            float acc = dot(nd.o1, ord1);
            float acc2 = 0.0f;
            for (int col = 0; col < 3; ++col) for (int row = 0; row < 3; ++row) {
                float h = (row == col ? inv : 0.0f) - (3.0f * rv[row]) * rv[col] / (4.0f * PI * r5);
                acc2 += nd.o2[col * 3 + row] * h;
            }
            return acc + acc2;
        }
        const Node& node = nodes[idx];
        if (node.leaf) {
            float wn = 0.0f;
            for (size_t t = node.left; t < node.right; ++t) wn += solid_angle(objects[t].first, p);
            if (COUNT) c->n_exact += node.right - node.left;
            return wn / (4.0f * 3.14159265358979323846f);
        }
        return fast_wn<COUNT>(node.left, p, beta, c) + fast_wn<COUNT>(node.right, p, beta, c);
    }
};

// ------------------------------------------------------------------------------------------------
struct ConvertStats {
    uint64_t n_tris, n_sub, n_eval, n_active, n_leaves, n_negative, wn_visit, wn_far, wn_exact, tree_nodes;
    double t_subdivide, t_tree, t_udf, t_sign;
};

inline idx_t f2i(float v) {  // Rust `as isize` (saturating; NaN -> 0)
    if (v != v) return 0;
    if (v >= 9.2e18f) return INT64_MAX;
    if (v <= -9.2e18f) return INT64_MIN;
    return idx_t(v);
}

template <class F> void parallel_for(size_t n, int threads, F f) {
    if (threads <= 1 || n < 2) { for (size_t i = 0; i < n; ++i) f(i); return; }
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back([&]() { for (;;) { size_t i = next.fetch_add(1); if (i >= n) break; f(i); } });
    for (auto& th : pool) th.join();
}

inline void subdivide_triangle(const Tri& tri, float voxel_size, std::vector<Tri>& out) {  // mesh_to_volume.rs:75-116
    float num_subs = std::floor(tri_max_side(tri) / voxel_size);
    float inv = 1.0f / num_subs;
    if (num_subs < 2.0f) { out.push_back(tri); return; }
    Vec3f s1 = (tri.b - tri.a) * inv, s2 = (tri.c - tri.b) * inv;
    Vec3f a = tri.a;
    size_t n = (num_subs != num_subs) ? 0 : (num_subs >= 1.8e19f ? SIZE_MAX : size_t(num_subs));  // Rust `as usize`
    for (size_t i = 0; i < n; ++i) {
        Vec3f b = a + s1, c = b + s2;
        Vec3f a_s = a + s2, b_s = b + s2, c_s = c + s2, a_prev = a;
        for (size_t k = 0; k < i; ++k) {
            out.push_back({a_prev, b_s, a_s});
            out.push_back({a_s, b_s, c_s});
            a_prev = a_s;
            a_s = a_s + s2; b_s = b_s + s2; c_s = c_s + s2;
        }
        out.push_back({a, b, c});
        a = a + s1;
    }
}

double now_s();

// MeshToVolume::convert (mesh_to_volume.rs:52-73). Returns nullptr where the reference returns None.
// count_work: also tally the winding-number traversal (per-leaf counter blocks, summed afterwards); off in every timed run
inline VolumeGrid* mesh_to_volume(const float* tris, size_t n_tris, float voxel_size, idx_t band, int threads, ConvertStats* st, bool count_work = false) {
    const float inverse_voxel_size = 1.0f / voxel_size;
    ConvertStats s; std::memset(&s, 0, sizeof(s));
    s.n_tris = n_tris;
    double t0 = now_s();
    std::vector<Tri> sub;
    for (size_t i = 0; i < n_tris; ++i) {
        const float* p = tris + 9 * i;
        subdivide_triangle(Tri{{p[0], p[1], p[2]}, {p[3], p[4], p[5]}, {p[6], p[7], p[8]}}, voxel_size, sub);
    }
    s.n_sub = sub.size();
    double t1 = now_s(); s.t_subdivide = t1 - t0;
    if (sub.empty()) { if (st) *st = s; return nullptr; }
    WindingNumbers wn; wn.build(tris, n_tris);
    s.tree_nodes = wn.nodes.size();
    double t2 = now_s(); s.t_tree = t2 - t1;

    // unsigned distance field (:118-196): parallel per sub-triangle evaluation, then serial strict-min merge
    VolumeGrid* grid = new VolumeGrid();
    struct Nb { Vec3i mn, mx; std::vector<float> d; };
    const size_t CHUNK = 1 << 16;
    std::vector<Nb> nb;
    for (size_t base = 0; base < sub.size(); base += CHUNK) {
        size_t cnt = std::min(CHUNK, sub.size() - base);
        nb.assign(cnt, Nb());
        const size_t BLK = 256, nblk = (cnt + BLK - 1) / BLK;
        parallel_for(nblk, threads, [&](size_t bi) {
            for (size_t j = bi * BLK; j < std::min(cnt, (bi + 1) * BLK); ++j) {
                const Tri& tri = sub[base + j];
                Box3f bb = tri_bbox(tri);
                Vec3i mn{f2i(std::floor(bb.mn.x * inverse_voxel_size)) - band, f2i(std::floor(bb.mn.y * inverse_voxel_size)) - band, f2i(std::floor(bb.mn.z * inverse_voxel_size)) - band};
                Vec3i mx{f2i(std::ceil(bb.mx.x * inverse_voxel_size)) + band, f2i(std::ceil(bb.mx.y * inverse_voxel_size)) + band, f2i(std::ceil(bb.mx.z * inverse_voxel_size)) + band};
                if (mx.x == mn.x || mx.y == mn.y || mx.z == mn.z) { mn = mn + Vec3i{-1, -1, -1}; mx = mx + Vec3i{1, 1, 1}; }
                Nb& o = nb[j]; o.mn = mn; o.mx = mx;
                for (idx_t x = mn.x; x <= mx.x; ++x) {
                    float xw = float(x) * voxel_size;
                    for (idx_t y = mn.y; y <= mx.y; ++y) {
                        float yw = float(y) * voxel_size;
                        for (idx_t z = mn.z; z <= mx.z; ++z) {
                            float zw = float(z) * voxel_size;
                            Vec3f gp{xw, yw, zw};
                            o.d.push_back(norm(closest_point(tri, gp) - gp));
                        }
                    }
                }
            }
        });
        for (size_t j = 0; j < cnt; ++j) {
            const Nb& o = nb[j];
            size_t i = 0;
            for (idx_t x = o.mn.x; x <= o.mx.x; ++x) for (idx_t y = o.mn.y; y <= o.mx.y; ++y) for (idx_t z = o.mn.z; z <= o.mx.z; ++z) {
                Vec3i idx{x, y, z};
                const float* cur = grid->at(idx);
                float cd = cur ? *cur : INFINITY;
                if (o.d[i] < cd) grid->insert(idx, o.d[i]);
                ++i;
            }
            s.n_eval += o.d.size();
        }
    }
    double t3 = now_s(); s.t_udf = t3 - t2;

    // signs (:198-281): same topology, value = copysign(|d|, wn < 0.2 ? + : -)
    struct Collect { std::vector<Leaf3<float>*> leaves; void dense(const Leaf3<float>& l) { leaves.push_back(const_cast<Leaf3<float>*>(&l)); } void tile(const Tile<float>&) {} } col;
    grid->visit_leafs(col);
    std::vector<WindingNumbers::Counters> wc(count_work ? col.leaves.size() : 0);
    parallel_for(col.leaves.size(), threads, [&](size_t li) {
        Leaf3<float>* leaf = col.leaves[li];
        WindingNumbers::Counters* cnt = count_work ? &wc[li] : nullptr;
        Vec3i o = leaf->origin();
        for (idx_t x = o.x; x < o.x + 8; ++x) for (idx_t y = o.y; y < o.y + 8; ++y) for (idx_t z = o.z; z < o.z + 8; ++z) {
            Vec3i idx{x, y, z};
            size_t off = Leaf3<float>::offset(idx);
            if (!leaf->value_mask.at(off)) continue;
            Vec3f gp{float(x) * voxel_size, float(y) * voxel_size, float(z) * voxel_size};
            float w = wn.approximate(gp, 2.0f, cnt);
            float d = leaf->values[off];
            leaf->values[off] = (w < 0.2f) ? std::copysign(d, 1.0f) : std::copysign(d, -1.0f);
        }
    });
    double t4 = now_s(); s.t_sign = t4 - t3;
    s.n_leaves = col.leaves.size();
    for (auto* l : col.leaves) for (int i = 0; i < 512; ++i) if (l->value_mask.at(i)) { s.n_active++; if (std::signbit(l->values[i])) s.n_negative++; }
    for (const auto& c : wc) { s.wn_visit += c.n_visit; s.wn_far += c.n_far; s.wn_exact += c.n_exact; }
    if (st) *st = s;
    return grid;
}

}  // namespace bso
