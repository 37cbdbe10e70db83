// Point-triangle distance down a z-column of lattice points, bit-identical to
// `(Triangle3::closest_point(p) - p).norm()` (src/geometry/primitives/triangle3.rs:317-382,
// src/voxel/mesh_to_volume.rs:155) and laid out for a warp whose lanes hold DIFFERENT triangles:
//
//  * nalgebra's dot is (x*x' + y*y') + z*z' with every operation rounded: along a z-column the (x*x' + y*y') part of all
//    six dots d1..d6 is the same for every point, so a point costs one multiply and one add per dot instead of five
//    operations plus the vector subtraction -- the same operations in the same order, so the same bits;
//  * the reference's seven Voronoi regions are a chain of early returns. Lanes of a warp land in different regions, and
//    a chain of branches would run every region's code one after the other. Here the region tests are evaluated as
//    predicates (first true wins, as in the chain), the three edge regions and the face region share ONE division
//    (numerator / denominator selected per lane) and the closest point is base + dir * q; only the face region has a
//    short tail of its own.
//
// The header compiles for the host as well (plain IEEE float operations; build with -ffp-contract=off) so that
// tests/test_ptdist_host.py can compare it bit for bit with a CPU restatement of the reference's closest_point (the
// checker under tests/host). The product only uses the device build.
#pragma once
#if defined(__CUDA_ARCH__)
#define PTD_FN __device__ __forceinline__
PTD_FN float ptd_add(float a, float b) { return __fadd_rn(a, b); }
PTD_FN float ptd_sub(float a, float b) { return __fsub_rn(a, b); }
PTD_FN float ptd_mul(float a, float b) { return __fmul_rn(a, b); }
PTD_FN float ptd_div(float a, float b) { return __fdiv_rn(a, b); }
PTD_FN float ptd_sqrt(float a) { return __fsqrt_rn(a); }
#else
#include <cmath>
#define PTD_FN static inline
PTD_FN float ptd_add(float a, float b) { return a + b; }
PTD_FN float ptd_sub(float a, float b) { return a - b; }
PTD_FN float ptd_mul(float a, float b) { return a * b; }
PTD_FN float ptd_div(float a, float b) { return a / b; }
PTD_FN float ptd_sqrt(float a) { return std::sqrt(a); }
#endif

struct PtdTri {  // vertices and the three edge vectors the reference forms: ab = b - a, ac = c - a, (c - b)
    float ax, ay, az, bx, by, bz, cx, cy, cz;
    float abx, aby, abz, acx, acy, acz, bcx, bcy, bcz;
};
struct PtdCol {  // column (x, y): world coordinates and the (x*x' + y*y') halves of d1..d6
    float xw, yw, s1, s2, s3, s4, s5, s6;
};

PTD_FN void ptd_tri_setup(PtdTri& T) {
    T.abx = ptd_sub(T.bx, T.ax); T.aby = ptd_sub(T.by, T.ay); T.abz = ptd_sub(T.bz, T.az);
    T.acx = ptd_sub(T.cx, T.ax); T.acy = ptd_sub(T.cy, T.ay); T.acz = ptd_sub(T.cz, T.az);
    T.bcx = ptd_sub(T.cx, T.bx); T.bcy = ptd_sub(T.cy, T.by); T.bcz = ptd_sub(T.cz, T.bz);
}
PTD_FN void ptd_col_setup(const PtdTri& T, float xw, float yw, PtdCol& K) {
    K.xw = xw; K.yw = yw;
    const float apx = ptd_sub(xw, T.ax), apy = ptd_sub(yw, T.ay);
    const float bpx = ptd_sub(xw, T.bx), bpy = ptd_sub(yw, T.by);
    const float cpx = ptd_sub(xw, T.cx), cpy = ptd_sub(yw, T.cy);
    K.s1 = ptd_add(ptd_mul(T.abx, apx), ptd_mul(T.aby, apy)); K.s2 = ptd_add(ptd_mul(T.acx, apx), ptd_mul(T.acy, apy));
    K.s3 = ptd_add(ptd_mul(T.abx, bpx), ptd_mul(T.aby, bpy)); K.s4 = ptd_add(ptd_mul(T.acx, bpx), ptd_mul(T.acy, bpy));
    K.s5 = ptd_add(ptd_mul(T.abx, cpx), ptd_mul(T.aby, cpy)); K.s6 = ptd_add(ptd_mul(T.acx, cpx), ptd_mul(T.acy, cpy));
}
// squared distance (the caller takes the root: sqrt is monotone, so it commutes with the scatter-min)
PTD_FN float ptd_eval2(const PtdTri& T, const PtdCol& K, float zw) {
    const float apz = ptd_sub(zw, T.az), bpz = ptd_sub(zw, T.bz), cqz = ptd_sub(zw, T.cz);
    const float d1 = ptd_add(K.s1, ptd_mul(T.abz, apz)), d2 = ptd_add(K.s2, ptd_mul(T.acz, apz));
    const float d3 = ptd_add(K.s3, ptd_mul(T.abz, bpz)), d4 = ptd_add(K.s4, ptd_mul(T.acz, bpz));
    const float d5 = ptd_add(K.s5, ptd_mul(T.abz, cqz)), d6 = ptd_add(K.s6, ptd_mul(T.acz, cqz));
    const float vc = ptd_sub(ptd_mul(d1, d4), ptd_mul(d3, d2));
    const float vb = ptd_sub(ptd_mul(d5, d2), ptd_mul(d1, d6));
    const float va = ptd_sub(ptd_mul(d3, d6), ptd_mul(d5, d4));
    const float e43 = ptd_sub(d4, d3), e56 = ptd_sub(d5, d6);
    // the reference's chain: A, B, AB, C, AC, BC, face -- the first true wins
    const bool tA = d1 <= 0.f && d2 <= 0.f;
    const bool tB = d3 >= 0.f && d4 <= d3;
    const bool tAB = vc <= 0.f && d1 >= 0.f && d3 <= 0.f;
    const bool tC = d6 >= 0.f && d5 <= d6;
    const bool tAC = vb <= 0.f && d2 >= 0.f && d6 <= 0.f;
    const bool tBC = va <= 0.f && e43 >= 0.f && e56 >= 0.f;
    const bool rA = tA, rB = !tA && tB, nAB_ = tA || tB;
    const bool rAB = !nAB_ && tAB, n3 = nAB_ || tAB;
    const bool rC = !n3 && tC, n4 = n3 || tC;
    const bool rAC = !n4 && tAC, n5 = n4 || tAC;
    const bool rBC = !n5 && tBC;
    const bool rF = !(n5 || tBC);
    const bool vertex = rA || rB || rC;
    // one division for the edge and face regions
    float num = 1.0f, den = ptd_add(ptd_add(va, vb), vc);                     // face: 1 / (va + vb + vc)
    if (rAB) { num = d1; den = ptd_sub(d1, d3); }                             // d1 / (d1 - d3)
    if (rAC) { num = d2; den = ptd_sub(d2, d6); }                             // d2 / (d2 - d6)
    if (rBC) { num = e43; den = ptd_add(e43, e56); }                          // (d4 - d3) / ((d4 - d3) + (d5 - d6))
    if (vertex) { num = 1.0f; den = 1.0f; }  // (not 0 / 1: a zero numerator sends the hardware division down its slow path)
    const float q = ptd_div(num, den);
    // base + dir * q: AB / AC start at a, BC at b
    float bx_ = T.ax, by_ = T.ay, bz_ = T.az, dx_ = T.abx, dy_ = T.aby, dz_ = T.abz;
    if (rAC) { dx_ = T.acx; dy_ = T.acy; dz_ = T.acz; }
    if (rBC || rB) { bx_ = T.bx; by_ = T.by; bz_ = T.bz; }
    if (rBC) { dx_ = T.bcx; dy_ = T.bcy; dz_ = T.bcz; }
    if (rC) { bx_ = T.cx; by_ = T.cy; bz_ = T.cz; }
    float px_ = ptd_add(bx_, ptd_mul(dx_, q)), py_ = ptd_add(by_, ptd_mul(dy_, q)), pz_ = ptd_add(bz_, ptd_mul(dz_, q));
    if (vertex) { px_ = bx_; py_ = by_; pz_ = bz_; }
    if (rF) {  // (a + ab * v) + ac * w with v = vb * denom, w = vc * denom
        const float v = ptd_mul(vb, q), w = ptd_mul(vc, q);
        px_ = ptd_add(ptd_add(T.ax, ptd_mul(T.abx, v)), ptd_mul(T.acx, w));
        py_ = ptd_add(ptd_add(T.ay, ptd_mul(T.aby, v)), ptd_mul(T.acy, w));
        pz_ = ptd_add(ptd_add(T.az, ptd_mul(T.abz, v)), ptd_mul(T.acz, w));
    }
    const float ex = ptd_sub(px_, K.xw), ey = ptd_sub(py_, K.yw), ez = ptd_sub(pz_, zw);
    return ptd_add(ptd_add(ptd_mul(ex, ex), ptd_mul(ey, ey)), ptd_mul(ez, ez));
}
