// The reference's own tests for this path, written against the C++ host mirror (include/baby_shark.hpp):
//   src/voxel/volume/mod.rs:134-152  box2.stl @0.2 -> offset(0.5) -> MC -> 7944 vertices
//   src/remeshing/voxel.rs:105-112   unit cube @0.1 -> faces > 0
//   examples/dual_contouring.rs      cuboid.subtract(sphere) -> DC gives a mesh
#include <cstdio>
#include <cstring>
#include "baby_shark.hpp"
using namespace baby_shark;

static const float BOX2[12 * 9] = {  // assets/box2.stl, file order
    -1, -1, 1, -1, 1, -1, -1, -1, -1,   -1, 1, 1, 1, 1, -1, -1, 1, -1,    1, 1, 1, 1, -1, -1, 1, 1, -1,   1, -1, 1, -1, -1, -1, 1, -1, -1,
    1, 1, -1, -1, -1, -1, -1, 1, -1,    -1, 1, 1, 1, -1, 1, 1, 1, 1,      -1, -1, 1, -1, 1, 1, -1, 1, -1,  -1, 1, 1, 1, 1, 1, 1, 1, -1,
    1, 1, 1, 1, -1, 1, 1, -1, -1,       1, -1, 1, -1, -1, 1, -1, -1, -1,  1, 1, -1, 1, -1, -1, -1, -1, -1, -1, 1, 1, -1, -1, 1, 1, -1, 1};

#define CHECK(c) do { if (!(c)) { std::printf("FAILED: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

int main() {
    {   // test_volume_offset
        auto volume = voxel::MeshToVolume().with_voxel_size(0.2f).convert(BOX2, 12);
        CHECK(volume.has_value());
        voxel::Volume off = std::move(*volume).offset(0.5f);
        auto vertices = voxel::MarchingCubesMesher().with_voxel_size(off.voxel_size()).mesh(off);
        CHECK(vertices.size() == 7944);
    }
    {   // test_voxel_remeshing (the cube as 12 triangles scaled to edge 1)
        float cube[12 * 9];
        for (int i = 0; i < 12 * 9; ++i) cube[i] = BOX2[i] * 0.5f;
        auto remeshed = remeshing::VoxelRemesher().with_voxel_size(0.1f).remesh(cube, 12);
        CHECK(remeshed.has_value() && remeshed->size() > 0 && remeshed->size() % 3 == 0);
        CHECK(!voxel::MeshToVolume().with_voxel_size(0.1f).convert(nullptr, 0).has_value());  // None on an empty mesh
    }
    {   // examples/dual_contouring.rs + consuming CSG
        auto builder = voxel::VolumeBuilder().with_voxel_size(0.2f);
        auto v = builder.cuboid({0, 0, 0}, {10, 10, 10}).subtract(builder.sphere(3.0f, {8, 8, 8}));
        auto tris = voxel::DualContouringMesher().with_voxel_size(0.2f).mesh(v);
        CHECK(tris.has_value() && tris->size() > 1000);
        bool panicked = false;  // union leaves active tiles -> the reference's DC hits todo!()
        try { auto u = builder.sphere(0.6f, {1.5f, 0.3f, 0.2f}).union_(builder.sphere(4.0f, {0.1f, 0.2f, 0.3f})); voxel::DualContouringMesher().with_voxel_size(0.2f).mesh(u); }
        catch (const Panic&) { panicked = true; }
        CHECK(panicked);
    }
    {   // io::stl + ActiveVoxelsMesher + merge_points on the unit-cube remesh
        std::vector<unsigned char> stl(84 + 12 * 50, 0);
        stl[80] = 12;
        for (int t = 0; t < 12; ++t) std::memcpy(&stl[84 + 50 * t + 12], BOX2 + 9 * t, 36);
        auto tris = io::StlReader().read_from_buffer(stl.data(), stl.size());
        CHECK(tris.size() == 12);
        voxel::MeshToVolume m2v;
        auto vol = voxel::convert(m2v, tris, 0.2f);
        CHECK(vol.has_value());
        auto soup = voxel::MarchingCubesMesher().with_voxel_size(0.2f).mesh(*vol);
        auto merged = algo::merge_points(soup);
        CHECK(merged.indices.size() == soup.size() && merged.points.size() * 3 < soup.size());
        CHECK((long)merged.points.size() - (long)soup.size() / 2 + (long)soup.size() / 3 == 2);  // closed genus-0 surface
        auto boxes = voxel::ActiveVoxelsMesher().mesh(*vol);
        CHECK(boxes.size() > 0 && boxes.size() % 6 == 0);
        bool failed = false;
        try { io::StlReader().read_from_buffer(stl.data(), stl.size() - 1); } catch (const Error&) { failed = true; }
        CHECK(failed);
    }
    std::printf("mirror ok\n");
    return 0;
}
