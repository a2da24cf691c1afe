// CPU check of the libfive::Heightmap stand-in that the reference drivers render_2d.cpp /
// render_3d.cpp link (mpr_b200/shim/src/heightmap_render.cpp): two spheres of radius 0.25 at
// x = +-0.5, as in those drivers' default shape.  Prints one line of numbers for the test.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <libfive/render/discrete/heightmap.hpp>
#include <libfive/tree/tree.hpp>

int main(int argc, char** argv) {
    auto X = libfive::Tree::X(), Y = libfive::Tree::Y(), Z = libfive::Tree::Z();
    auto t = min(sqrt((X + 0.5) * (X + 0.5) + Y * Y + Z * Z) - 0.25,
                 sqrt((X - 0.5) * (X - 0.5) + Y * Y + Z * Z) - 0.25);
    std::atomic_bool abort(false);
    const int n = 64;
    auto h = libfive::Heightmap::render(t, libfive::Voxels({-1, -1, -1}, {1, 1, 1}, n / 2), abort);
    auto g = libfive::Heightmap::render(t, libfive::Voxels({-1, -1, 0}, {1, 1, 0}, n / 2), abort);
    int filled3 = 0, filled2 = 0, wrong = 0;
    float zmax = -1e9f;
    for (int y = 0; y < n; ++y)
        for (int x = 0; x < n; ++x) {
            const float fx = -1 + (x + 0.5f) * 2 / n, fy = -1 + (y + 0.5f) * 2 / n;
            const float d = std::min(std::hypot(fx + 0.5f, fy), std::hypot(fx - 0.5f, fy)) - 0.25f;
            const bool in3 = h->depth(y, x) > -1e9f, in2 = g->depth(y, x) > -1e9f;
            filled3 += in3;
            filled2 += in2;
            if (in3) zmax = std::max(zmax, h->depth(y, x));
            // the silhouette of a sphere is its equatorial disc (up to one voxel of z sampling)
            if (in2 != (d < 0)) ++wrong;
            if (in3 && d > 0.05f) ++wrong;
            if (!in3 && d < -0.05f) ++wrong;
        }
    const uint32_t top = h->norm(n / 2, n / 4);      // above the centre of the left sphere: normal = +z
    if (argc > 1) { h->savePNG(std::string(argv[1]) + "/depth.png"); h->saveNormalPNG(std::string(argv[1]) + "/norm.png"); }
    printf("%d %d %d %.6f %08x\n", filled3, filled2, wrong, zmax, top);
    return 0;
}
