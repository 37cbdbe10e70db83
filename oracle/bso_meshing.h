// ORACLE -- TEST INFRASTRUCTURE ONLY (see bso_tree.h). CPU restatement of
//   src/voxel/meshing/marching_cubes.rs:43-63 (mesh), :72-285 (handle_cube), :287-334 (add_faces,
//   intersection), :336-376 (interior_test_case13), :378-401 (test_face), :403-472 (test_interior),
//   :474-538 (interior_ambiguity), :540-916 (interior_ambiguity_verification), :918-938
//   (compute_c_vertex), :970-1010 (CubesVisitor), :1013-1121 (ComputeEdgeIntersections), :1161-1181;
//   src/voxel/meshing/lookup_table.rs (tables; flattened by tools/gen_mc33_tables.py);
//   src/voxel/meshing/dual_contouring.rs:23-83,93-387,389-458; src/voxel/utils.rs:86-95.
#pragma once
#include "bso_ops.h"
#include "../baby_shark_b200/csrc/mc33_tables.h"

namespace bso {

static const int8_t MC33[MC33_BLOB_SIZE] = MC33_BLOB_INIT;
static const uint8_t IAV_PERM[12][8] = MC33_IAV_PERM_INIT;
static const int8_t EDGE_V1[13] = {0, 1, 3, 0, 4, 5, 7, 4, 0, 1, 2, 3, 0};
static const int8_t EDGE_V2[13] = {1, 2, 2, 3, 5, 6, 6, 7, 4, 5, 6, 7, 0};
static const int8_t EDGE_DIR[12] = {0, 1, 0, 1, 0, 1, 0, 1, 2, 2, 2, 2};  // X,Y,Z per lookup_table.rs:45-52
static const Vec3i CUBE_OFFSETS[8] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
static const float MIN_ABS_VERTEX_VALUE = 1e-6f;

#define T1(name, cfg) (&MC33[MC33_OFF_##name + (cfg) * MC33_ROW_##name]), MC33_ROW_##name
#define T2(name, cfg, sub) (&MC33[MC33_OFF_##name + ((cfg) * MC33_SUB_##name + (sub)) * MC33_ROW_##name]), MC33_ROW_##name
#define TEST1(name, cfg) (MC33[MC33_OFF_##name + (cfg)])
#define TEST2(name, cfg, i) (MC33[MC33_OFF_##name + (cfg) * MC33_ROW_##name + (i)])

struct McStats { uint64_t case_hist[15]; uint64_t n_cubes; uint64_t n_degenerate; };

struct MarchingCubes {
    const VolumeGrid* grid;
    VolumeGrid x_int, y_int, z_int;
    std::vector<Vec3f> vertices;
    float voxel_size;
    Vec3f v12{0, 0, 0};
    struct Cube { uint8_t id; Vec3i index[8]; float v[8]; } cube;
    int mc_case = 0, config = 0;
    McStats stats;

    float c(int i) const { return cube.v[i]; }

    // ---- pass 1 (:1013-1121)
    void compute_intersection(const Vec3i& v1, float a, float b, int dir) {
        if (sign_of(a) == sign_of(b)) return;
        a = std::fmax(std::fabs(a), MIN_ABS_VERTEX_VALUE);
        b = std::fmax(std::fabs(b), MIN_ABS_VERTEX_VALUE);
        float t = a / (a + b);
        if (dir == 0) x_int.insert(v1, float(v1.x) + t);
        else if (dir == 1) y_int.insert(v1, float(v1.y) + t);
        else z_int.insert(v1, float(v1.z) + t);
    }
    struct Pass1 {
        MarchingCubes* mc;
        void dense(const Leaf3<float>& l) {
            Vec3i o = l.origin();
            for (idx_t x = o.x; x < o.x + 8; ++x) for (idx_t y = o.y; y < o.y + 8; ++y) for (idx_t z = o.z; z < o.z + 8; ++z) {
                Vec3i v{x, y, z};
                const float* a = mc->grid->at(v);
                if (!a) continue;
                const Vec3i nb[3] = {{x + 1, y, z}, {x, y + 1, z}, {x, y, z + 1}};
                for (int d = 0; d < 3; ++d) { const float* b = mc->grid->at(nb[d]); if (b) mc->compute_intersection(v, *a, *b, d); }
            }
        }
        void tile(const Tile<float>& t) {
            idx_t s = idx_t(t.size);
            for (idx_t i = 0; i < s; ++i) for (idx_t j = 0; j < s; ++j) {
                Vec3i right = t.origin + Vec3i{s - 1, i, j}, front = t.origin + Vec3i{i, s - 1, j}, top = t.origin + Vec3i{i, j, s - 1};
                const float* b;
                if ((b = mc->grid->at(right + Vec3i{1, 0, 0}))) mc->compute_intersection(right, t.value, *b, 0);
                if ((b = mc->grid->at(front + Vec3i{0, 1, 0}))) mc->compute_intersection(front, t.value, *b, 1);
                if ((b = mc->grid->at(top + Vec3i{0, 0, 1}))) mc->compute_intersection(top, t.value, *b, 2);
            }
        }
    };
    // ---- pass 2
    bool cube_from_voxel(const Vec3i& voxel) {  // :1161-1181
        cube.id = 0;
        for (int i = 0; i < 8; ++i) {
            Vec3i idx = voxel + CUBE_OFFSETS[i];
            const float* p = grid->at(idx);
            if (!p) return false;
            float v = *p;
            if (std::fabs(v) < MIN_ABS_VERTEX_VALUE) v = std::copysign(MIN_ABS_VERTEX_VALUE, v);
            if (v < 0.0f) cube.id |= uint8_t(1 << i);
            cube.index[i] = idx; cube.v[i] = v;
        }
        return true;
    }
    struct Pass2 {
        MarchingCubes* mc;
        void dense(const Leaf3<float>& l) {
            Vec3i o = l.origin();
            for (idx_t x = o.x; x < o.x + 8; ++x) for (idx_t y = o.y; y < o.y + 8; ++y) for (idx_t z = o.z; z < o.z + 8; ++z) mc->handle_voxel(Vec3i{x, y, z});
        }
        void tile(const Tile<float>& t) {  // boundary voxels only; edge/corner voxels are visited more than once, as in the reference
            Vec3i o = t.origin; idx_t s = idx_t(t.size);
            for (idx_t i = 0; i < s; ++i) for (idx_t j = 0; j < s; ++j) {
                mc->handle_voxel(o + Vec3i{0, i, j});      // left
                mc->handle_voxel(o + Vec3i{s - 1, i, j});  // right
                mc->handle_voxel(o + Vec3i{i, j, s - 1});  // top
                mc->handle_voxel(o + Vec3i{i, j, 0});      // bottom
                mc->handle_voxel(o + Vec3i{i, s - 1, j});  // front
                mc->handle_voxel(o + Vec3i{i, 0, j});      // back
            }
        }
    };
    void handle_voxel(const Vec3i& v) { if (cube_from_voxel(v)) { stats.n_cubes++; handle_cube(); } }

    bool intersection(int e, Vec3f& out) const {  // :320-334
        if (e == 12) { out = v12; return true; }
        const Vec3i& idx = cube.index[EDGE_V1[e]];
        const float* p;
        switch (EDGE_DIR[e]) {
            case 0: p = x_int.at(idx); if (!p) return false; out = Vec3f{*p, float(idx.y), float(idx.z)}; return true;
            case 1: p = y_int.at(idx); if (!p) return false; out = Vec3f{float(idx.x), *p, float(idx.z)}; return true;
            default: p = z_int.at(idx); if (!p) return false; out = Vec3f{float(idx.x), float(idx.y), *p}; return true;
        }
    }
    void add_faces(const int8_t* edges, int n) {  // :287-317 ; emission order (e[i], e[i+2], e[i+1])
        for (int i = 0; i + 2 < n; i += 3) {
            int e1 = edges[i], e3 = edges[i + 1], e2 = edges[i + 2];
            Vec3f v1, v2, v3;
            if (!intersection(e1, v1) || !intersection(e2, v2) || !intersection(e3, v3)) continue;
            v1 = v1 * voxel_size; v2 = v2 * voxel_size; v3 = v3 * voxel_size;
            if (tri_is_degenerate(v1, v2, v3)) { stats.n_degenerate++; continue; }
            vertices.push_back(v1); vertices.push_back(v2); vertices.push_back(v3);
        }
    }
    void compute_c_vertex() {  // :918-938
        Vec3f sum{0, 0, 0}; int count = 0;
        for (int e = 0; e < 12; ++e) { Vec3f p; if (intersection(e, p)) { sum = sum + p; ++count; } }
        v12 = sum / float(count);
    }
    bool test_face(int face) const {  // :378-401
        float a, b, cc, d;
        switch (face) {
            case -1: case 1: a = c(0); b = c(4); cc = c(5); d = c(1); break;
            case -2: case 2: a = c(1); b = c(5); cc = c(6); d = c(2); break;
            case -3: case 3: a = c(2); b = c(6); cc = c(7); d = c(3); break;
            case -4: case 4: a = c(3); b = c(7); cc = c(4); d = c(0); break;
            case -5: case 5: a = c(0); b = c(3); cc = c(2); d = c(1); break;
            case -6: case 6: a = c(4); b = c(7); cc = c(6); d = c(5); break;
            default: return false;
        }
        float val = a * cc - b * d;
        if (std::fabs(val) < FLT_EPSILON) return face >= 0;
        return float(face) * a * val >= 0.0f;
    }
    int interior_ambiguity(int amb_face, int face_i) const {  // :474-538
        float face = float(face_i);
        int edge = 0;
        auto pos = [&](int i, int j) { return c(i) * face > 0.0f && c(j) * face > 0.0f; };
        switch (amb_face) {
            case 1: case 3:
                if (pos(1, 7)) edge = 4;
                if (pos(0, 6)) edge = 5;
                if (pos(3, 5)) edge = 6;
                if (pos(2, 4)) edge = 7;
                break;
            case 2: case 4:
                if (pos(1, 7)) edge = 0;
                if (pos(2, 4)) edge = 1;
                if (pos(3, 5)) edge = 2;
                if (pos(0, 6)) edge = 3;
                break;
            case 5: case 6: case 0:
                if (pos(0, 6)) edge = 8;
                if (pos(1, 7)) edge = 9;
                if (pos(2, 4)) edge = 10;
                if (pos(3, 5)) edge = 11;
                break;
            default: break;
        }
        return edge;
    }
    int interior_ambiguity_verification(int edge) const {  // :540-916, one formula under IAV_PERM
        if (edge < 0 || edge > 11) return 0;
        const uint8_t* p = IAV_PERM[edge];
        float A0 = c(p[0]), A1 = c(p[1]), B0 = c(p[2]), B1 = c(p[3]), C0 = c(p[4]), C1 = c(p[5]), D0 = c(p[6]), D1 = c(p[7]);
        float a = (A1 - A0) * (C1 - C0) - (B1 - B0) * (D1 - D0);
        float b = C0 * (A1 - A0) + A0 * (C1 - C0) - D0 * (B1 - B0) - B0 * (D1 - D0);
        if (a > 0.0f) return 1;
        float t = -b / (2.0f * a);
        if (t < 0.0f || t > 1.0f) return 1;
        float at = A0 + (A1 - A0) * t, bt = B0 + (B1 - B0) * t, ct = C0 + (C1 - C0) * t, dt = D0 + (D1 - D0) * t;
        float verify = at * ct - bt * dt;
        if (verify > 0.0f) return 0;
        if (verify < 0.0f) return 1;
        return 0;
    }
    bool test_interior(int face) const {  // :403-472
        switch (mc_case) {
            case 4: {
                int amb = interior_ambiguity_verification(interior_ambiguity(1, face));
                amb += interior_ambiguity_verification(interior_ambiguity(2, face));
                amb += interior_ambiguity_verification(interior_ambiguity(5, face));
                return amb != 0;
            }
            case 6: return interior_ambiguity_verification(interior_ambiguity(std::abs(int(TEST2(TEST_6, config, 0))), face)) != 0;
            case 7: {
                int s = -face;
                int amb = interior_ambiguity_verification(interior_ambiguity(1, s));
                amb += interior_ambiguity_verification(interior_ambiguity(2, s));
                amb += interior_ambiguity_verification(interior_ambiguity(5, s));
                return amb != 0;
            }
            case 10: return interior_ambiguity_verification(interior_ambiguity(std::abs(int(TEST2(TEST_10, config, 0))), face)) != 0;
            case 12: {
                int amb = interior_ambiguity_verification(interior_ambiguity(std::abs(int(TEST2(TEST_12, config, 0))), face));
                amb += interior_ambiguity_verification(interior_ambiguity(std::abs(int(TEST2(TEST_12, config, 1))), face));
                return amb != 0;
            }
            default: return false;
        }
    }
    bool interior_test_case13() const {  // :336-376
        float a = (c(0) - c(1)) * (c(7) - c(6)) - (c(4) - c(5)) * (c(3) - c(2));
        float b = c(6) * (c(0) - c(1)) + c(1) * (c(7) - c(6)) - c(2) * (c(4) - c(5)) - c(5) * (c(3) - c(2));
        float cc = c(1) * c(6) - c(5) * c(2);
        float delta = b * b - 4.0f * a * cc;
        float t1 = (-b + std::sqrt(delta)) / (a + a), t2 = (-b - std::sqrt(delta)) / (a + a);
        if (t1 < 1.0f && t1 > 0.0f && t2 < 1.0f && t2 > 0.0f) {
            float a1 = c(1) + (c(0) - c(1)) * t1, b1 = c(5) + (c(4) - c(5)) * t1, c1 = c(6) + (c(7) - c(6)) * t1, d1 = c(2) + (c(3) - c(2)) * t1;
            float x1 = (a1 - d1) / (a1 + c1 - b1 - d1), y1 = (a1 - b1) / (a1 + c1 - b1 - d1);
            float a2 = c(1) + (c(0) - c(1)) * t2, b2 = c(5) + (c(4) - c(5)) * t2, c2 = c(6) + (c(7) - c(6)) * t2, d2 = c(2) + (c(3) - c(2)) * t2;
            float x2 = (a2 - d2) / (a2 + c2 - b2 - d2), y2 = (a2 - b2) / (a2 + c2 - b2 - d2);
            return !(x1 < 1.0f && x1 > 0.0f && x2 < 1.0f && x2 > 0.0f && y1 < 1.0f && y1 > 0.0f && y2 < 1.0f && y2 > 0.0f);
        }
        return true;
    }
    void handle_cube() {  // :72-285
        mc_case = MC33[MC33_OFF_CASES + 2 * cube.id]; config = MC33[MC33_OFF_CASES + 2 * cube.id + 1];
        if (mc_case >= 0 && mc_case < 15) stats.case_hist[mc_case]++;
        const int cf = config;
        switch (mc_case) {
            case 0: break;
            case 1: add_faces(T1(TILING_1, cf)); break;
            case 2: add_faces(T1(TILING_2, cf)); break;
            case 3: if (test_face(TEST1(TEST_3, cf))) add_faces(T1(TILING_3_2, cf)); else add_faces(T1(TILING_3_1, cf)); break;
            case 4: if (test_interior(TEST1(TEST_4, cf))) add_faces(T1(TILING_4_1, cf)); else add_faces(T1(TILING_4_2, cf)); break;
            case 5: add_faces(T1(TILING_5, cf)); break;
            case 6:
                if (test_face(TEST2(TEST_6, cf, 0))) add_faces(T1(TILING_6_2, cf));
                else if (test_interior(TEST2(TEST_6, cf, 1))) add_faces(T1(TILING_6_1_1, cf));
                else add_faces(T1(TILING_6_1_2, cf));
                break;
            case 7: {
                int sub = 0;
                if (test_face(TEST2(TEST_7, cf, 0))) sub += 1;
                if (test_face(TEST2(TEST_7, cf, 1))) sub += 2;
                if (test_face(TEST2(TEST_7, cf, 2))) sub += 4;
                switch (sub) {
                    case 0: add_faces(T1(TILING_7_1, cf)); break;
                    case 1: add_faces(T2(TILING_7_2, cf, 0)); break;
                    case 2: add_faces(T2(TILING_7_2, cf, 1)); break;
                    case 3: compute_c_vertex(); add_faces(T2(TILING_7_3, cf, 0)); break;
                    case 4: add_faces(T2(TILING_7_2, cf, 2)); break;
                    case 5: compute_c_vertex(); add_faces(T2(TILING_7_3, cf, 1)); break;
                    case 6: compute_c_vertex(); add_faces(T2(TILING_7_3, cf, 2)); break;
                    case 7: if (test_interior(TEST2(TEST_7, cf, 3))) add_faces(T1(TILING_7_4_1, cf)); else add_faces(T1(TILING_7_4_2, cf)); break;
                }
                break;
            }
            case 8: add_faces(T1(TILING_8, cf)); break;
            case 9: add_faces(T1(TILING_9, cf)); break;
            case 10:
                if (test_face(TEST2(TEST_10, cf, 0))) {
                    if (test_face(TEST2(TEST_10, cf, 1))) {
                        if (test_interior(-TEST2(TEST_10, cf, 2))) add_faces(T1(TILING_10_1_1_, cf)); else add_faces(T1(TILING_10_1_2, 5 - cf));
                    } else { compute_c_vertex(); add_faces(T1(TILING_10_2, cf)); }
                } else if (test_face(TEST2(TEST_10, cf, 1))) { compute_c_vertex(); add_faces(T1(TILING_10_2_, cf)); }
                else if (test_interior(TEST2(TEST_10, cf, 2))) add_faces(T1(TILING_10_1_1, cf));
                else add_faces(T1(TILING_10_1_2, cf));
                break;
            case 11: add_faces(T1(TILING_11, cf)); break;
            case 12:
                if (test_face(TEST2(TEST_12, cf, 0))) {
                    if (test_face(TEST2(TEST_12, cf, 1))) {
                        if (test_interior(-TEST2(TEST_12, cf, 2))) add_faces(T1(TILING_12_1_1_, cf)); else add_faces(T1(TILING_12_1_2, 23 - cf));
                    } else { compute_c_vertex(); add_faces(T1(TILING_12_2, cf)); }
                } else if (test_face(TEST2(TEST_12, cf, 1))) { compute_c_vertex(); add_faces(T1(TILING_12_2_, cf)); }
                else if (test_interior(TEST2(TEST_12, cf, 2))) add_faces(T1(TILING_12_1_1, cf));
                else add_faces(T1(TILING_12_1_2, cf));
                break;
            case 13: {
                int sub = 0;
                for (int i = 0; i < 6; ++i) if (test_face(TEST2(TEST_13, cf, i))) sub += 1 << i;
                int sc = MC33[MC33_OFF_SUB_CONFIG_13 + sub];
                if (sc == 0) add_faces(T1(TILING_13_1, cf));
                else if (sc >= 1 && sc <= 6) add_faces(T2(TILING_13_2, cf, sc - 1));
                else if (sc >= 7 && sc <= 18) { compute_c_vertex(); add_faces(T2(TILING_13_3, cf, sc - 7)); }
                else if (sc >= 19 && sc <= 22) { compute_c_vertex(); add_faces(T2(TILING_13_4, cf, sc - 19)); }
                else if (sc >= 23 && sc <= 26) {
                    if (cf == 0) { if (interior_test_case13()) add_faces(T2(TILING_13_5_1, 0, sc - 23)); else add_faces(T2(TILING_13_5_2, 1, sc - 23)); }
                    else if (interior_test_case13()) add_faces(T2(TILING_13_5_1, 1, sc - 23));
                    else add_faces(T2(TILING_13_5_2, 0, sc - 23));
                }
                else if (sc >= 27 && sc <= 38) { compute_c_vertex(); add_faces(T2(TILING_13_3_, cf, sc - 27)); }
                else if (sc >= 39 && sc <= 44) add_faces(T2(TILING_13_2_, cf, sc - 39));
                else if (sc == 45) add_faces(T1(TILING_13_1_, cf));
                break;
            }
            case 14: add_faces(T1(TILING_14, cf)); break;
            default: break;
        }
    }
    void mesh(const Volume& vol, float vs) {
        grid = vol.grid; voxel_size = vs; vertices.clear();
        std::memset(&stats, 0, sizeof(stats));
        Pass1 p1{this}; grid->visit_leafs(p1);
        Pass2 p2{this}; grid->visit_leafs(p2);
    }
};

// ------------------------------------------------------------------------------------------------
// Dual contouring (dual_contouring.rs)
struct IntPoint { Vec3f point, normal; };
inline bool operator==(const IntPoint& a, const IntPoint& b) { return a.point == b.point && a.normal == b.normal; }
struct DualContouringError { const char* what; };

struct DualContouring {
    const VolumeGrid* grid;
    Grid<IntPoint> xi, yi, zi;
    Grid<Vec3f> cells;
    std::vector<Vec3f> faces, out;

    float grad(const Vec3i& p, int axis) const {  // :326-342
        Vec3i pl = p, pr = p;
        (axis == 0 ? pl.x : axis == 1 ? pl.y : pl.z) -= 1;
        (axis == 0 ? pr.x : axis == 1 ? pr.y : pr.z) += 1;
        const float *vp = grid->at(p), *vl = grid->at(pl), *vr = grid->at(pr);
        if (vp && vl && vr) return (*vr - *vl) * 0.5f;
        if (vp && vr) return *vr - *vp;
        if (vp && vl) return *vp - *vl;
        throw DualContouringError{"unreachable!(): gradient has no neighbour (dual_contouring.rs:340)"};
    }
    void intersection(const Vec3i& v1, int dir) {  // :266-307
        Vec3i v2 = v1; (dir == 0 ? v2.x : dir == 1 ? v2.y : v2.z) += 1;
        const float *a = grid->at(v1), *b = grid->at(v2);
        if (!a || !b) return;
        if (sign_of(*a) == sign_of(*b)) return;
        float t = (*a == *b) ? 0.5f : *a / (*a - *b);
        Vec3f point{float(v1.x), float(v1.y), float(v1.z)};
        (dir == 0 ? point.x : dir == 1 ? point.y : point.z) += t;
        float g[3];
        for (int ax = 0; ax < 3; ++ax) g[ax] = (1.0f - t) * grad(v1, ax) + t * grad(v2, ax);
        Vec3f n{g[0], g[1], g[2]};
        n = n / norm(n);
        (dir == 0 ? xi : dir == 1 ? yi : zi).insert(v1, IntPoint{point, n});
    }
    static Vec3f find_feature_point(const std::vector<IntPoint>& pts) {  // :429-458
        const float threshold = 1e-6f; const int iters = 50;
        Vec3f c{0, 0, 0};
        for (auto& p : pts) c = c + p.point;
        c = c / float(pts.size());
        for (int i = 0; i < iters; ++i) {
            Vec3f force{0, 0, 0};
            for (auto& ip : pts) force = force + (ip.normal * -1.0f) * dot(ip.normal, c - ip.point);
            float damping = 1.0f - float(i) / float(iters);
            c = c + (force * damping) / float(pts.size());
            if (norm_squared(force) < threshold) break;
        }
        return c;
    }
    struct P1 { DualContouring* dc; void tile(const Tile<float>&) { throw DualContouringError{"todo!(): tile support"}; }
        void dense(const Leaf3<float>& l) { Vec3i o = l.origin();
            for (idx_t x = o.x; x < o.x + 8; ++x) for (idx_t y = o.y; y < o.y + 8; ++y) for (idx_t z = o.z; z < o.z + 8; ++z) { Vec3i v{x, y, z}; dc->intersection(v, 0); dc->intersection(v, 1); dc->intersection(v, 2); } } };
    struct P2 { DualContouring* dc; void tile(const Tile<float>&) { throw DualContouringError{"todo!(): tile support"}; }
        void dense(const Leaf3<float>& l) {
            static const int EO[12][4] = {{0,0,0,0},{0,0,0,1},{0,0,0,2},{1,0,0,1},{1,0,0,2},{0,0,1,0},{0,0,1,1},{0,1,0,0},{0,1,0,2},{1,0,1,1},{0,1,1,0},{1,1,0,2}};
            Vec3i o = l.origin(); std::vector<IntPoint> ints;
            for (idx_t x = o.x; x < o.x + 8; ++x) for (idx_t y = o.y; y < o.y + 8; ++y) for (idx_t z = o.z; z < o.z + 8; ++z) {
                Vec3i q{x, y, z}; float vals[8]; bool ok = true;
                for (int i = 0; i < 8 && ok; ++i) { const float* p = dc->grid->at(q + CUBE_OFFSETS[i]); if (!p) ok = false; else vals[i] = *p; }
                if (!ok) continue;
                bool same = true; for (int i = 1; i < 8; ++i) if (sign_of(vals[i]) != sign_of(vals[0])) same = false;
                if (same) continue;
                ints.clear();
                for (int e = 0; e < 12; ++e) { Vec3i p = q + Vec3i{EO[e][0], EO[e][1], EO[e][2]}; const IntPoint* ip = (EO[e][3] == 0 ? dc->xi : EO[e][3] == 1 ? dc->yi : dc->zi).at(p); if (ip) ints.push_back(*ip); }
                dc->cells.insert(q, find_feature_point(ints));
            } } };
    struct P3 { DualContouring* dc; void tile(const Tile<float>&) { throw DualContouringError{"todo!(): tile support"}; }
        void handle_edge(float v1_val, const Vec3i& v1, int dir) {  // :99-135
            static const int CO[3][4][3] = {{{0,0,0},{0,0,-1},{0,-1,-1},{0,-1,0}}, {{0,0,0},{-1,0,0},{-1,0,-1},{0,0,-1}}, {{0,-1,0},{-1,-1,0},{-1,0,0},{0,0,0}}};
            Vec3i v2 = v1; (dir == 0 ? v2.x : dir == 1 ? v2.y : v2.z) += 1;
            const float* b = dc->grid->at(v2);
            if (!b) return;
            if (sign_of(v1_val) == sign_of(*b)) return;
            const Vec3f* p[4];
            for (int i = 0; i < 4; ++i) { p[i] = dc->cells.at(v1 + Vec3i{CO[dir][i][0], CO[dir][i][1], CO[dir][i][2]}); if (!p[i]) return; }
            Vec3f f[6] = {*p[0], *p[1], *p[2], *p[2], *p[3], *p[0]};
            if (sign_of(v1_val) == Negative) { std::swap(f[1], f[2]); std::swap(f[4], f[5]); }
            for (int i = 0; i < 6; ++i) dc->faces.push_back(f[i]);
        }
        void dense(const Leaf3<float>& l) { Vec3i o = l.origin();
            for (idx_t x = o.x; x < o.x + 8; ++x) for (idx_t y = o.y; y < o.y + 8; ++y) for (idx_t z = o.z; z < o.z + 8; ++z) {
                Vec3i v{x, y, z}; const float* a = l.at(v); if (!a) continue;
                handle_edge(*a, v, 0); handle_edge(*a, v, 1); handle_edge(*a, v, 2); } } };
    // returns false where the reference panics / returns None
    bool mesh(const Volume& vol, float vs, const char** err) {
        grid = vol.grid; faces.clear(); out.clear();
        try {
            P1 p1{this}; grid->visit_leafs(p1);
            P2 p2{this}; grid->visit_leafs(p2);
            P3 p3{this}; grid->visit_leafs(p3);
        } catch (DualContouringError& e) { if (err) *err = e.what; return false; }
        for (size_t i = 0; i + 2 < faces.size(); i += 3) {
            Vec3f v0 = faces[i] * vs, v1 = faces[i + 1] * vs, v2 = faces[i + 2] * vs;
            if (tri_is_degenerate(v0, v1, v2)) continue;
            out.push_back(v0); out.push_back(v1); out.push_back(v2);
        }
        return true;
    }
};

}  // namespace bso
