// Stand-in for libfive's CPU renderer, Heightmap::render (libfive/libfive/src/render/discrete/
// heightmap.cpp:255-317), which the reference drivers render_2d.cpp / render_3d.cpp call to write
// their out_cpu.png comparison image.  It is linked into those drivers only (never into
// libmprb.so): the product has no CPU render path.
//
// Same contract as libfive's: the result has rows = y samples, cols = x samples; depth(y, x) is
// the z of the highest voxel centre where the expression is negative (-inf when there is none),
// norm(y, x) the packed unit gradient there (0xffff7f7f on the top face).  The method is a plain
// region recursion over an interval evaluation of the expression DAG, with point samples at
// single voxels; it makes no attempt to reproduce libfive's tape shortening.
#include <algorithm>
#include <cmath>
#include <limits>
#include <map>
#include <thread>
#include <vector>

#include "libfive/render/discrete/heightmap.hpp"

namespace libfive {
namespace {

struct Flat {
    struct N { Opcode::Opcode op; int a, b; float v; };
    std::vector<N> nodes;       // operands before users; the last node is the root
    explicit Flat(const Tree& t) {
        std::map<Tree::Id, int> index;
        for (const Tree& n : t.orderedDfs()) {
            N f{n->op, -1, -1, n->value};
            if (n->lhs) f.a = index.at(n->lhs.get());
            if (n->rhs) f.b = index.at(n->rhs.get());
            index[n.id()] = int(nodes.size());
            nodes.push_back(f);
        }
    }
};

inline float dn(float v) { return std::nextafter(v, -std::numeric_limits<float>::infinity()); }
inline float up(float v) { return std::nextafter(v, std::numeric_limits<float>::infinity()); }
struct I { float lo, hi; };
const I kAny = {-std::numeric_limits<float>::infinity(), std::numeric_limits<float>::infinity()};

I imul(I a, I b) {
    const float c[4] = {a.lo * b.lo, a.lo * b.hi, a.hi * b.lo, a.hi * b.hi};
    float lo = c[0], hi = c[0];
    for (float x : c) { if (!(x == x)) return kAny; lo = std::min(lo, x); hi = std::max(hi, x); }
    return {dn(lo), up(hi)};
}
template <typename F> I mono(I a, F f) { return {dn(f(a.lo)), up(f(a.hi))}; }

// Conservative interval of every node; anything not handled exactly widens to (-inf, inf).
void eval_interval(const Flat& f, I x, I y, I z, std::vector<I>& out) {
    using namespace Opcode;
    out.resize(f.nodes.size());
    for (size_t i = 0; i < f.nodes.size(); ++i) {
        const Flat::N& n = f.nodes[i];
        const I a = n.a >= 0 ? out[n.a] : kAny, b = n.b >= 0 ? out[n.b] : kAny;
        I r = kAny;
        switch (n.op) {
            case CONSTANT: r = {n.v, n.v}; break;
            case VAR_X: r = x; break;
            case VAR_Y: r = y; break;
            case VAR_Z: r = z; break;
            case OP_SQUARE: {
                const float l = a.lo * a.lo, h = a.hi * a.hi;
                r = (a.lo <= 0 && a.hi >= 0) ? I{0.0f, up(std::max(l, h))} : I{dn(std::min(l, h)), up(std::max(l, h))};
                break;
            }
            case OP_SQRT: r = a.hi < 0 ? kAny : I{a.lo <= 0 ? 0.0f : dn(std::sqrt(a.lo)), up(std::sqrt(a.hi))}; break;
            case OP_NEG: r = {-a.hi, -a.lo}; break;
            case OP_SIN: case OP_COS: r = {-1.0f, 1.0f}; break;
            case OP_ASIN: if (a.lo >= -1 && a.hi <= 1) r = mono(a, [](float v) { return std::asin(v); }); break;
            case OP_ACOS: if (a.lo >= -1 && a.hi <= 1) { r = {dn(std::acos(a.hi)), up(std::acos(a.lo))}; } break;
            case OP_ATAN: r = mono(a, [](float v) { return std::atan(v); }); break;
            case OP_EXP: r = mono(a, [](float v) { return std::exp(v); }); break;
            case OP_LOG: if (a.lo > 0) r = mono(a, [](float v) { return std::log(v); }); break;
            case OP_ABS:
                r = a.lo >= 0 ? a : a.hi <= 0 ? I{-a.hi, -a.lo} : I{0.0f, std::max(-a.lo, a.hi)};
                break;
            case OP_RECIP: if (a.lo > 0 || a.hi < 0) r = {dn(1.0f / a.hi), up(1.0f / a.lo)}; break;
            case OP_ADD: r = {dn(a.lo + b.lo), up(a.hi + b.hi)}; break;
            case OP_SUB: r = {dn(a.lo - b.hi), up(a.hi - b.lo)}; break;
            case OP_MUL: r = imul(a, b); break;
            case OP_DIV: if (b.lo > 0 || b.hi < 0) r = imul(a, I{dn(1.0f / b.hi), up(1.0f / b.lo)}); break;
            case OP_MIN: r = {std::min(a.lo, b.lo), std::min(a.hi, b.hi)}; break;
            case OP_MAX: r = {std::max(a.lo, b.lo), std::max(a.hi, b.hi)}; break;
            case CONST_VAR: r = a; break;
            default: break;
        }
        if (!(r.lo == r.lo) || !(r.hi == r.hi)) r = kAny;
        out[i] = r;
    }
}

struct D { float v, dx, dy, dz; };

// Value and gradient at a point (forward mode).
D eval_point(const Flat& f, float x, float y, float z, std::vector<D>& s) {
    using namespace Opcode;
    s.resize(f.nodes.size());
    for (size_t i = 0; i < f.nodes.size(); ++i) {
        const Flat::N& n = f.nodes[i];
        const D a = n.a >= 0 ? s[n.a] : D{0, 0, 0, 0}, b = n.b >= 0 ? s[n.b] : D{0, 0, 0, 0};
        auto chain = [&](float v, float k) { return D{v, a.dx * k, a.dy * k, a.dz * k}; };
        D r{std::nanf(""), 0, 0, 0};
        switch (n.op) {
            case CONSTANT: r = {n.v, 0, 0, 0}; break;
            case VAR_X: r = {x, 1, 0, 0}; break;
            case VAR_Y: r = {y, 0, 1, 0}; break;
            case VAR_Z: r = {z, 0, 0, 1}; break;
            case OP_SQUARE: r = chain(a.v * a.v, 2 * a.v); break;
            case OP_SQRT: r = chain(std::sqrt(a.v), a.v > 0 ? 0.5f / std::sqrt(a.v) : 0.0f); break;
            case OP_NEG: r = chain(-a.v, -1); break;
            case OP_SIN: r = chain(std::sin(a.v), std::cos(a.v)); break;
            case OP_COS: r = chain(std::cos(a.v), -std::sin(a.v)); break;
            case OP_TAN: r = chain(std::tan(a.v), 1 / (std::cos(a.v) * std::cos(a.v))); break;
            case OP_ASIN: r = chain(std::asin(a.v), 1 / std::sqrt(1 - a.v * a.v)); break;
            case OP_ACOS: r = chain(std::acos(a.v), -1 / std::sqrt(1 - a.v * a.v)); break;
            case OP_ATAN: r = chain(std::atan(a.v), 1 / (1 + a.v * a.v)); break;
            case OP_EXP: r = chain(std::exp(a.v), std::exp(a.v)); break;
            case OP_LOG: r = chain(std::log(a.v), 1 / a.v); break;
            case OP_ABS: r = chain(std::fabs(a.v), a.v < 0 ? -1.0f : 1.0f); break;
            case OP_RECIP: r = chain(1 / a.v, -1 / (a.v * a.v)); break;
            case CONST_VAR: r = a; break;
            case OP_ADD: r = {a.v + b.v, a.dx + b.dx, a.dy + b.dy, a.dz + b.dz}; break;
            case OP_SUB: r = {a.v - b.v, a.dx - b.dx, a.dy - b.dy, a.dz - b.dz}; break;
            case OP_MUL:
                r = {a.v * b.v, a.dx * b.v + b.dx * a.v, a.dy * b.v + b.dy * a.v, a.dz * b.v + b.dz * a.v};
                break;
            case OP_DIV: {
                const float q = b.v * b.v;
                r = {a.v / b.v, (a.dx * b.v - b.dx * a.v) / q, (a.dy * b.v - b.dy * a.v) / q,
                     (a.dz * b.v - b.dz * a.v) / q};
                break;
            }
            case OP_MIN: r = a.v < b.v ? a : b; break;
            case OP_MAX: r = a.v < b.v ? b : a; break;
            case OP_ATAN2: {
                const float q = a.v * a.v + b.v * b.v;
                r = {std::atan2(a.v, b.v), (a.dx * b.v - b.dx * a.v) / q, (a.dy * b.v - b.dy * a.v) / q,
                     (a.dz * b.v - b.dz * a.v) / q};
                break;
            }
            case OP_POW: {
                const float p = std::pow(a.v, b.v), k = b.v * std::pow(a.v, b.v - 1);
                r = {p, a.dx * k, a.dy * k, a.dz * k};
                break;
            }
            default: break;
        }
        s[i] = r;
    }
    return s.back();
}

struct Job {
    const Flat& f;
    const Voxels& vox;
    Heightmap& out;
    const std::atomic_bool& abort;
    std::vector<I> ivals;
    std::vector<D> pvals;

    uint32_t pack_normal(const D& d) const {
        float n = std::sqrt(d.dx * d.dx + d.dy * d.dy + d.dz * d.dz);
        if (!(n > 0)) n = 1;
        const int nx = int(255 * (d.dx / n / 2 + 0.5f)), ny = int(255 * (d.dy / n / 2 + 0.5f)),
                  nz = int(255 * (d.dz / n / 2 + 0.5f));
        return (0xffu << 24) | (uint32_t(nz & 0xff) << 16) | (uint32_t(ny & 0xff) << 8) | uint32_t(nx & 0xff);
    }
    I span(int axis, int lo, int hi) const {      // hull of sample positions lo..hi-1
        const auto& p = vox.pts[axis];
        return {p[lo], p[hi - 1]};
    }
    void fill(int x0, int x1, int y0, int y1, int ztop) {
        const float z = vox.pts[2][ztop];
        for (int y = y0; y < y1; ++y)
            for (int x = x0; x < x1; ++x)
                if (out.depth(y, x) < z) {
                    out.depth(y, x) = z;
                    out.norm(y, x) = pack_normal(eval_point(f, vox.pts[0][x], vox.pts[1][y], z, pvals));
                }
    }
    void recurse(int x0, int x1, int y0, int y1, int z0, int z1) {
        if (abort.load()) return;
        // nothing to do if every pixel already holds something at or above this region's top
        const float ztop = vox.pts[2][z1 - 1];
        bool open = false;
        for (int y = y0; y < y1 && !open; ++y)
            for (int x = x0; x < x1; ++x)
                if (out.depth(y, x) < ztop) { open = true; break; }
        if (!open) return;
        const int nx = x1 - x0, ny = y1 - y0, nz = z1 - z0;
        if (nx == 1 && ny == 1 && nz == 1) {
            const D d = eval_point(f, vox.pts[0][x0], vox.pts[1][y0], vox.pts[2][z0], pvals);
            if (d.v < 0) {
                out.depth(y0, x0) = vox.pts[2][z0];
                out.norm(y0, x0) = pack_normal(d);
            }
            return;
        }
        eval_interval(f, span(0, x0, x1), span(1, y0, y1), span(2, z0, z1), ivals);
        const I r = ivals.back();
        if (r.lo > 0) return;
        if (r.hi < 0) { fill(x0, x1, y0, y1, z1 - 1); return; }
        // split the longest axis; the upper z half goes first so that it can hide the lower one
        if (nz >= nx && nz >= ny) {
            const int m = z0 + nz / 2;
            recurse(x0, x1, y0, y1, m, z1);
            recurse(x0, x1, y0, y1, z0, m);
        } else if (nx >= ny) {
            const int m = x0 + nx / 2;
            recurse(x0, m, y0, y1, z0, z1);
            recurse(m, x1, y0, y1, z0, z1);
        } else {
            const int m = y0 + ny / 2;
            recurse(x0, x1, y0, m, z0, z1);
            recurse(x0, x1, m, y1, z0, z1);
        }
    }
};

}  // namespace

std::unique_ptr<Heightmap> Heightmap::render(const Tree t, Voxels r, const std::atomic_bool& abort, size_t threads) {
    const int nx = int(r.pts[0].size()), ny = int(r.pts[1].size()), nz = int(r.pts[2].size());
    std::unique_ptr<Heightmap> out(new Heightmap(ny, nx));
    out->depth = -std::numeric_limits<float>::infinity();
    const Flat flat(t);
    // bands of rows, one per worker: disjoint pixels, so no synchronisation is needed
    const int workers = int(std::max<size_t>(1, std::min<size_t>(threads, size_t(ny))));
    std::vector<std::thread> pool;
    for (int w = 0; w < workers; ++w) {
        const int y0 = ny * w / workers, y1 = ny * (w + 1) / workers;
        pool.emplace_back([&, y0, y1]() {
            Job job{flat, r, *out, abort, {}, {}};
            if (y1 > y0) job.recurse(0, nx, y0, y1, 0, nz);
        });
    }
    for (auto& th : pool) th.join();
    if (nz > 0) {
        const float top = r.pts[2].back();
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x)
                if (out->depth(y, x) == top) out->norm(y, x) = 0xffff7f7fu;
    }
    return out;
}

}  // namespace libfive
