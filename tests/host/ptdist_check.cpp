// Host-side check of baby_shark_b200/csrc/bs_ptdist.cuh (the column formulation of the point-triangle distance used by
// k_eval) against the oracle's closest_point (oracle/bso_convert.h, which restates triangle3.rs:317-382): bit equality of
// the distances on random and degenerate triangles. Test infrastructure; built and run by tests/test_ptdist_host.py with
// -ffp-contract=off.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include "../../oracle/bso_tree.h"
#include "../../oracle/bso_convert.h"
#include "../../baby_shark_b200/csrc/bs_ptdist.cuh"

static uint32_t bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

int main(int argc, char** argv) {
    const long n_tri = argc > 1 ? atol(argv[1]) : 200000;
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<float> U(-1.0f, 1.0f);
    long bad = 0, total = 0, region_hist[2] = {0, 0};
    for (long i = 0; i < n_tri; ++i) {
        const int kind = (int)(i % 8);
        const float vs = kind == 7 ? 0.37f : 1.0f / 2048.0f;
        bso::Vec3f c0{U(rng), U(rng), U(rng)};
        auto jitter = [&](float s) { return bso::Vec3f{U(rng) * s, U(rng) * s, U(rng) * s}; };
        bso::Tri t;
        const float s = vs * (kind == 6 ? 8.0f : 1.5f);
        t.a = c0 + jitter(s); t.b = c0 + jitter(s); t.c = c0 + jitter(s);
        if (kind == 1) { t.c = t.a + (t.b - t.a) * 0.37f; }                       // collinear (up to rounding)
        if (kind == 2) { t.c = t.b; }                                             // duplicate vertex
        if (kind == 3) { t.b = t.a; t.c = t.a; }                                  // a point
        if (kind == 4) { t.c = t.a + (t.b - t.a) * 0.5f + jitter(vs * 1e-5f); }   // sliver
        if (kind == 5) {                                                          // lattice-aligned vertices (ties in the region tests)
            auto snap = [&](float v) { return std::floor(v / vs) * vs; };
            t.a = {snap(t.a.x), snap(t.a.y), snap(t.a.z)}; t.b = {snap(t.b.x), snap(t.b.y), snap(t.b.z)}; t.c = {snap(t.c.x), snap(t.c.y), snap(t.c.z)};
        }
        PtdTri T{t.a.x, t.a.y, t.a.z, t.b.x, t.b.y, t.b.z, t.c.x, t.c.y, t.c.z, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        ptd_tri_setup(T);
        const int x0 = (int)std::floor(c0.x / vs) - 2, y0 = (int)std::floor(c0.y / vs) - 2, z0 = (int)std::floor(c0.z / vs) - 2;
        for (int xi = 0; xi < 5; ++xi)
            for (int yi = 0; yi < 5; ++yi) {
                const float xw = (float)(x0 + xi) * vs, yw = (float)(y0 + yi) * vs;
                PtdCol K;
                ptd_col_setup(T, xw, yw, K);
                for (int zi = 0; zi < 5; ++zi) {
                    const float zw = (float)(z0 + zi) * vs;
                    const bso::Vec3f p{xw, yw, zw};
                    const float ref = bso::norm(bso::closest_point(t, p) - p);
                    const float got = ptd_sqrt(ptd_eval2(T, K, zw));
                    ++total;
                    if (bits(ref) != bits(got) && !(ref != ref && got != got)) {
                        if (bad < 10) std::printf("mismatch kind %d: ref %.9g (%08x) got %.9g (%08x)\n", kind, ref, bits(ref), got, bits(got));
                        ++bad;
                    }
                    region_hist[ref == 0.0f]++;
                }
            }
    }
    std::printf("checked %ld distances, %ld mismatches, %ld exact zeros\n", total, bad, region_hist[1]);
    return bad ? 1 : 0;
}
