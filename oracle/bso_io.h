// ORACLE -- TEST INFRASTRUCTURE ONLY (see bs_oracle.cpp). CPU restatement of the data formats either side of the
// path: binary STL reader / writer, ActiveVoxelsMesher, merge_points.
#pragma once
#include "bso_meshing.h"
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>

namespace bso {

// io/stl.rs:65-95 (StlReader::read_from_buffer): 80-byte header, u32 triangle count, then 50-byte records
// {normal 3 f32 (ignored), v1, v2, v3, u16 attribute}. Returns false on a short buffer (read_exact -> ReadError).
inline bool stl_decode(const uint8_t* bytes, size_t n_bytes, std::vector<float>& tris) {
    if (n_bytes < 84) return false;
    uint32_t n; std::memcpy(&n, bytes + 80, 4);
    if (n_bytes < 84 + (size_t)n * 50) return false;
    tris.resize((size_t)n * 9);
    for (size_t i = 0; i < n; ++i) std::memcpy(tris.data() + 9 * i, bytes + 84 + 50 * i + 12, 36);
    return true;
}

// io/stl.rs:143-191 (StlWriter::write_to_buffer) for a triangle soup: zero header, count, per face the recomputed
// normal (Triangle3::normal, triangle3.rs:261-269: cross.normalize(), zeros for a degenerate face), vertices, 0u16.
inline void stl_encode(const float* verts, size_t n_tris, std::vector<uint8_t>& out) {
    out.assign(84 + n_tris * 50, 0);
    const uint32_t n = (uint32_t)n_tris;
    std::memcpy(out.data() + 80, &n, 4);
    for (size_t i = 0; i < n_tris; ++i) {
        const float* p = verts + 9 * i;
        const Vec3f a{p[0], p[1], p[2]}, b{p[3], p[4], p[5]}, c{p[6], p[7], p[8]};
        const Vec3f cr = cross(b - a, c - a);
        Vec3f nrm{0.0f, 0.0f, 0.0f};
        if (!(norm_squared(cr) == 0.0f)) nrm = cr / norm(cr);
        uint8_t* r = out.data() + 84 + 50 * i;
        std::memcpy(r, &nrm, 12); std::memcpy(r + 12, p, 36);
    }
}

// voxel/meshing/active_voxels.rs:12-126: every active voxel (leaf visit order, x outer / z inner; tiles: boundary
// voxels in the order left,right,top,bottom,front,back per (i,j), duplicates included) emits two triangles for each
// of its six faces whose neighbour is not active, faces in the order top(+z) bottom(-z) left(-x) right(+x) front(+y) back(-y).
struct ActiveVoxels {
    const VolumeGrid* grid;
    std::vector<Vec3i> out;
    void test_voxel(const Vec3i& v) {
        if (!grid->at(v)) return;
        static const Vec3i B[8] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};  // utils.rs:86-95
        static const int F[6][6] = {{4, 6, 7, 4, 5, 6}, {1, 0, 3, 1, 3, 2}, {0, 4, 3, 4, 7, 3}, {1, 6, 5, 1, 2, 6}, {2, 3, 6, 6, 3, 7}, {1, 5, 0, 5, 4, 0}};
        static const Vec3i N[6] = {{0, 0, 1}, {0, 0, -1}, {-1, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, -1, 0}};
        bool open[6];
        for (int f = 0; f < 6; ++f) open[f] = !grid->at(v + N[f]);  // all six neighbours are probed before anything is emitted
        for (int f = 0; f < 6; ++f) if (open[f]) for (int k = 0; k < 6; ++k) out.push_back(v + B[F[f][k]]);
    }
    void dense(const Leaf3<float>& l) {
        const Vec3i o = l.origin();
        for (idx_t x = o.x; x < o.x + 8; ++x) for (idx_t y = o.y; y < o.y + 8; ++y) for (idx_t z = o.z; z < o.z + 8; ++z) test_voxel(Vec3i{x, y, z});
    }
    void tile(const Tile<float>& t) {
        const Vec3i o = t.origin; const idx_t s = idx_t(t.size);
        for (idx_t i = 0; i < s; ++i) for (idx_t j = 0; j < s; ++j) {
            test_voxel(o + Vec3i{0, i, j}); test_voxel(o + Vec3i{s - 1, i, j});
            test_voxel(o + Vec3i{i, j, s - 1}); test_voxel(o + Vec3i{i, j, 0});
            test_voxel(o + Vec3i{i, s - 1, j}); test_voxel(o + Vec3i{i, 0, j});
        }
    }
};

// algo/merge_points.rs:12-41 with data_structures/vertex_index_map.rs: exactly coincident points (f32 ==, so +0 == -0
// and NaN never equals anything) share an index; unique points keep first-occurrence order.
struct PointKey { float x, y, z; };
struct PointKeyHash { size_t operator()(const PointKey& k) const { uint32_t a, b, c; float x = k.x + 0.0f, y = k.y + 0.0f, z = k.z + 0.0f; std::memcpy(&a, &x, 4); std::memcpy(&b, &y, 4); std::memcpy(&c, &z, 4); return ((size_t)a * 73856093u) ^ ((size_t)b * 19349663u) ^ ((size_t)c * 83492791u); } };
struct PointKeyEq { bool operator()(const PointKey& a, const PointKey& b) const { return a.x == b.x && a.y == b.y && a.z == b.z; } };
inline void merge_points(const float* pts, size_t n, std::vector<float>& unique, std::vector<uint32_t>& indices) {
    std::unordered_map<PointKey, uint32_t, PointKeyHash, PointKeyEq> map;
    map.reserve(n / 3 + 1);
    indices.resize(n); unique.clear();
    for (size_t i = 0; i < n; ++i) {
        const PointKey k{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
        auto it = map.find(k);
        if (it != map.end()) { indices[i] = it->second; continue; }
        const uint32_t id = (uint32_t)(unique.size() / 3);
        unique.insert(unique.end(), {k.x, k.y, k.z});
        map.emplace(k, id);  // a NaN key can be inserted but never found again: every NaN point stays unique, as in the reference
        indices[i] = id;
    }
}

}  // namespace bso
