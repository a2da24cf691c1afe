// Minimal libfive::Heightmap (include/libfive/render/discrete/heightmap.hpp:19-76): the
// depth / normal image container the drivers fill from mpr::Context results and save.
// savePNG / saveNormalPNG write real PNG files (stored-deflate, no libpng needed).
// Heightmap::render - libfive's CPU renderer - is a declared-only stand-in here: the CPU
// comparison leg of this repository lives in oracle/ (test infrastructure), not in the product.
#pragma once
#include <atomic>
#include <memory>
#include <string>
#include <Eigen/Eigen>

#include "libfive/tree/tree.hpp"
#include "libfive/render/discrete/voxels.hpp"

namespace libfive {
class Heightmap {
public:
    Heightmap(unsigned rows, unsigned cols) : depth(rows, cols), norm(rows, cols) { depth = 0.0f; norm = 0u; }
    static std::unique_ptr<Heightmap> render(const Tree t, Voxels r, const std::atomic_bool& abort,
                                             size_t threads = 8);
    bool savePNG(std::string filename);         // 16-bit grey, depth scaled to its range
    bool saveNormalPNG(std::string filename);   // 8-bit RGBA from the packed normals
    typedef Eigen::Array<float, Eigen::Dynamic, Eigen::Dynamic> Depth;
    typedef Eigen::Array<uint32_t, Eigen::Dynamic, Eigen::Dynamic> Normal;
    Depth depth;
    Normal norm;
};
}  // namespace libfive
