// TEST / BENCH INFRASTRUCTURE ONLY -- not part of the product path (libmprb.so has no CPU render path).
//
// C entry point around the libfive-algorithm CPU renderer that the unchanged reference drivers link
// (mpr_b200/shim/src/heightmap_render.cpp, a stand-in for libfive's Heightmap::render,
// libfive/libfive/src/render/discrete/heightmap.cpp:195-318 - libfive itself needs Eigen, Boost and libpng,
// none of which exist here), so that bench.py can time it on the host cores next to the GPU arms:
// BASELINE.json's "libfive's own CPU renderer timed on the box's host cores ... as a reported baseline".
//
// Input is the packed clause tape every other arm consumes (the .frep models do not travel to the GPU
// box); the expression is rebuilt from it clause by clause - slots name the sub-expressions, exactly
// the inverse of src/tape.cpp:111-214 - and rendered as the drivers do it: render_2d.cpp:72-74
// (Voxels({-1,-1,0},{1,1,0}, S/2)) and render_3d.cpp:82-84 (Voxels({-1,-1,-1},{1,1,1}, S/2)).
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>

#include <libfive/render/discrete/heightmap.hpp>
#include <libfive/tree/tree.hpp>

using libfive::Tree;

static Tree tree_from_tape(const uint64_t* cells, int32_t n) {
    std::vector<Tree> slot(256, Tree(0.0f));
    const uint32_t h = uint32_t(cells[0]);
    if ((h >> 8) & 0xff) slot[(h >> 8) & 0xff] = Tree::X();
    if ((h >> 16) & 0xff) slot[(h >> 16) & 0xff] = Tree::Y();
    if (h >> 24) slot[h >> 24] = Tree::Z();
    for (int32_t i = 1; i + 1 < n; ++i) {
        const uint32_t w = uint32_t(cells[i]);
        const uint32_t op = w & 0xff, out = (w >> 8) & 0xff;
        const Tree L = slot[(w >> 16) & 0xff], R = slot[w >> 24];
        float imm;
        const uint32_t hi = uint32_t(cells[i] >> 32);
        memcpy(&imm, &hi, 4);
        const Tree I(imm);
        Tree r(0.0f);
        switch (op) {                  // reference inc/gpu_opcode.hpp:18-56
            case 2: r = square(L); break;
            case 3: r = sqrt(L); break;
            case 4: r = -L; break;
            case 5: r = sin(L); break;
            case 6: r = cos(L); break;
            case 7: r = asin(L); break;
            case 8: r = acos(L); break;
            case 9: r = atan(L); break;
            case 10: r = exp(L); break;
            case 11: r = abs(L); break;
            case 12: r = log(L); break;
            case 13: r = L + I; break;
            case 14: r = L + R; break;
            case 15: r = L * I; break;
            case 16: r = L * R; break;
            case 17: r = min(L, I); break;
            case 18: r = min(L, R); break;
            case 19: r = max(L, I); break;
            case 20: r = max(L, R); break;
            case 21: r = L - I; break;
            case 22: r = I - R; break;
            case 23: r = L - R; break;
            case 24: r = L / I; break;
            case 25: r = I / R; break;
            case 26: r = L / R; break;
            case 27: r = I; break;
            case 28: r = L; break;
            case 29: r = R; break;
            default: break;
        }
        slot[out] = r;
    }
    return slot[(uint32_t(cells[n - 1]) >> 8) & 0xff];
}

extern "C" {

// Renders one frame; returns the milliseconds Heightmap::render took (evaluator construction
// included, as in the drivers' calls).  depth_out (size*size floats) may be null.
double mpro_heightmap_ms(const uint64_t* cells, int32_t n, int dim, int size, int threads, float* depth_out) {
    const Tree t = tree_from_tape(cells, n);
    std::atomic_bool abort(false);
    const libfive::Voxels vox = dim == 3 ? libfive::Voxels({-1, -1, -1}, {1, 1, 1}, size / 2.0f)
                                         : libfive::Voxels({-1, -1, 0}, {1, 1, 0}, size / 2.0f);
    const auto t0 = std::chrono::steady_clock::now();
    auto h = libfive::Heightmap::render(t, vox, abort, size_t(threads > 0 ? threads : 8));
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (depth_out)
        for (int y = 0; y < size; ++y)
            for (int x = 0; x < size; ++x) depth_out[size_t(y) * size + x] = h->depth(y, x);
    return ms;
}

}  // extern "C"
