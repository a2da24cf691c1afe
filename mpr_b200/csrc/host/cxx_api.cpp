// Out-of-line pieces of the mpr:: C++ surface (mpr_b200/inc/mpr.hpp) and of the libfive
// stand-ins that need a translation unit: Tape construction from a Tree, opcode names,
// Heightmap PNG output.
#include <cstring>
#include <fstream>
#include <limits>
#include <vector>

#include "libfive/render/discrete/heightmap.hpp"
#include "libfive/tree/tree.hpp"
#include "mpr.hpp"
#include "mprb_host.hpp"

namespace mpr {

Tape::Tape(const libfive::Tree& tree) {
    int n_slots = 0;
    const std::vector<uint64_t> cells = mprb::pack_tape(tree, &n_slots);
    mprb_tape* t = nullptr;
    detail::check(mprb_tape_create(cells.data(), int32_t(cells.size()), &t), "mprb_tape_create");
    handle.reset(t);
    data = detail::borrowed<uint64_t[]>(const_cast<uint64_t*>(mprb_tape_data(t)));
    length = mprb_tape_length(t);
}

Tape::Tape(const uint64_t* cells, int32_t n) {
    mprb_tape* t = nullptr;
    detail::check(mprb_tape_create(cells, n, &t), "mprb_tape_create");
    handle.reset(t);
    data = detail::borrowed<uint64_t[]>(const_cast<uint64_t*>(mprb_tape_data(t)));
    length = mprb_tape_length(t);
}

const char* gpu_op_str(uint8_t op) {
    static const char* names[] = {
        "INVALID", "JUMP", "SQUARE_LHS", "SQRT_LHS", "NEG_LHS", "SIN_LHS", "COS_LHS", "ASIN_LHS", "ACOS_LHS",
        "ATAN_LHS", "EXP_LHS", "ABS_LHS", "LOG_LHS", "ADD_LHS_IMM", "ADD_LHS_RHS", "MUL_LHS_IMM", "MUL_LHS_RHS",
        "MIN_LHS_IMM", "MIN_LHS_RHS", "MAX_LHS_IMM", "MAX_LHS_RHS", "SUB_LHS_IMM", "SUB_IMM_RHS", "SUB_LHS_RHS",
        "DIV_LHS_IMM", "DIV_IMM_RHS", "DIV_LHS_RHS", "COPY_IMM", "COPY_LHS", "COPY_RHS"};
    return op < sizeof(names) / sizeof(names[0]) ? names[op] : "UNKNOWN";
}

}  // namespace mpr

////////////////////////////////////////////////////////////////////////////////
// PNG output without libpng: stored (uncompressed) deflate blocks inside a zlib stream.

namespace {

uint32_t crc32_of(const uint8_t* p, size_t n, uint32_t crc = 0) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    return ~crc;
}

void put32(std::vector<uint8_t>& v, uint32_t x) { for (int s = 24; s >= 0; s -= 8) v.push_back(uint8_t(x >> s)); }

void chunk(std::ofstream& f, const char* type, const std::vector<uint8_t>& body) {
    std::vector<uint8_t> buf(type, type + 4);
    buf.insert(buf.end(), body.begin(), body.end());
    std::vector<uint8_t> len;
    put32(len, uint32_t(body.size()));
    f.write(reinterpret_cast<const char*>(len.data()), 4);
    f.write(reinterpret_cast<const char*>(buf.data()), std::streamsize(buf.size()));
    std::vector<uint8_t> crc;
    put32(crc, crc32_of(buf.data(), buf.size()));
    f.write(reinterpret_cast<const char*>(crc.data()), 4);
}

bool write_png(const std::string& name, unsigned w, unsigned h, int color_type, int depth,
               const std::vector<uint8_t>& rows /* h * (1 + stride) bytes, filter byte included */) {
    std::ofstream f(name, std::ios::binary);
    if (!f.is_open()) return false;
    const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    f.write(reinterpret_cast<const char*>(sig), 8);
    std::vector<uint8_t> ihdr;
    put32(ihdr, w);
    put32(ihdr, h);
    ihdr.push_back(uint8_t(depth));
    ihdr.push_back(uint8_t(color_type));
    ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk(f, "IHDR", ihdr);
    std::vector<uint8_t> z = {0x78, 0x01};
    uint32_t a = 1, b = 0;
    for (uint8_t c : rows) { a = (a + c) % 65521; b = (b + a) % 65521; }
    for (size_t off = 0; off < rows.size() || off == 0; off += 65535) {
        const size_t n = std::min<size_t>(65535, rows.size() - off);
        z.push_back(off + n >= rows.size() ? 1 : 0);
        z.push_back(uint8_t(n)); z.push_back(uint8_t(n >> 8));
        z.push_back(uint8_t(~n)); z.push_back(uint8_t((~n) >> 8));
        z.insert(z.end(), rows.begin() + long(off), rows.begin() + long(off + n));
        if (rows.empty()) break;
    }
    put32(z, (b << 16) | a);
    chunk(f, "IDAT", z);
    chunk(f, "IEND", {});
    return bool(f);
}

}  // namespace

namespace libfive {

// Same pixel mapping as libfive (src/render/discrete/heightmap.cpp:378-399): empty (-inf) pixels
// are black, the finite range maps to 1..65535, and the image is written bottom row first.
bool Heightmap::savePNG(std::string filename) {
    const unsigned h = unsigned(depth.rows()), w = unsigned(depth.cols());
    const float ninf = -std::numeric_limits<float>::infinity();
    const float zmax = depth.maxCoeff();
    float zmin = zmax;
    for (unsigned r = 0; r < h; ++r)
        for (unsigned c = 0; c < w; ++c)
            if (depth(r, c) != ninf && depth(r, c) < zmin) zmin = depth(r, c);
    std::vector<uint8_t> rows;
    rows.reserve(size_t(h) * (1 + 2 * size_t(w)));
    for (unsigned k = 0; k < h; ++k) {
        const unsigned r = h - 1 - k;
        rows.push_back(0);
        for (unsigned c = 0; c < w; ++c) {
            const float d = depth(r, c);
            const float sc = (zmax == zmin) ? (d - zmin) + 65535.0f : (d - zmin) * 65534.0f / (zmax - zmin) + 1.0f;
            const unsigned v = (d == ninf || !(sc > 0.0f)) ? 0u : (sc >= 65535.0f ? 65535u : unsigned(sc));
            rows.push_back(uint8_t(v >> 8));
            rows.push_back(uint8_t(v));
        }
    }
    return write_png(filename, w, h, 0, 16, rows);
}

bool Heightmap::saveNormalPNG(std::string filename) {
    const unsigned h = unsigned(norm.rows()), w = unsigned(norm.cols());
    std::vector<uint8_t> rows;
    rows.reserve(size_t(h) * (1 + 4 * size_t(w)));
    for (unsigned k = 0; k < h; ++k) {
        const unsigned r = h - 1 - k;
        rows.push_back(0);
        for (unsigned c = 0; c < w; ++c) {
            const uint32_t p = norm(r, c);
            rows.push_back(uint8_t(p)); rows.push_back(uint8_t(p >> 8));
            rows.push_back(uint8_t(p >> 16)); rows.push_back(uint8_t(p >> 24));
        }
    }
    return write_png(filename, w, h, 6, 8, rows);
}

}  // namespace libfive
