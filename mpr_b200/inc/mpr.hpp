// mpr:: C++ surface over the mprb C ABI (include/mprb.h).
//
// Re-creates, member for member, what the reference's drivers use from inc/tape.hpp,
// inc/context.hpp, inc/util.hpp, inc/clause.hpp, inc/gpu_opcode.hpp and inc/parameters.hpp
// (SURVEY.md section 8b lists every use site), so that benchmark/*.cpp compile unchanged
// against this header set and link against libmprb.so.  Everything forwards to the C ABI;
// there is no rendering code in this header.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <memory>

#include <Eigen/Eigen>

#include "../../include/mprb.h"

namespace libfive { class Tree; }

////////////////////////////////////////////////////////////////////////////////
// inc/clause.hpp: clause field accessors (d points at a 64-bit clause)
#define OP(d) (((uint8_t*)(d))[0])
#define I_OUT(d) (((uint8_t*)(d))[1])
#define I_LHS(d) (((uint8_t*)(d))[2])
#define I_RHS(d) (((uint8_t*)(d))[3])
#define IMM(d) (((float*)(d))[1])
#define JUMP_TARGET(d) (((int32_t*)(d))[1])

// inc/parameters.hpp
#ifndef NUM_TILES
#define NUM_TILES (4)
#define NUM_THREADS (64 * NUM_TILES)
#define SUBTAPE_CHUNK_SIZE 64
#ifdef BIG_SERVER
#define NUM_SUBTAPES 6400000
#else
#define NUM_SUBTAPES 640000
#endif
#endif

// inc/util.hpp:19-41.  CUDA_MALLOC / CUDA_FREE go through the C ABI, so host-only drivers need no
// CUDA toolkit; CUDA_CHECK exists only where the CUDA runtime header does (brute.cu, gui/tex.cu).
#ifndef CUDA_MALLOC
namespace mpr { namespace detail {
template <typename T> inline T* managed_checked(size_t count, const char* file, int line) {
    void* p = nullptr;
    if (mprb_malloc_managed(sizeof(T) * count, &p) != 0) {
        fprintf(stderr, "Error: %s %s %d\n", mprb_last_error(), file, line);
        exit(1);
    }
    return static_cast<T*>(p);
}
} }
#define CUDA_MALLOC(T, c) mpr::detail::managed_checked<T>(c, __FILE__, __LINE__)
#define CUDA_FREE(c) mprb_free_device((void*)(c))
#endif
#if defined(__CUDACC__) && !defined(CUDA_CHECK)
#define CUDA_CHECK(f) { mpr_gpu_check((f), __FILE__, __LINE__); }
inline void mpr_gpu_check(cudaError_t code, const char* file, int line) {
    if (code != cudaSuccess) {
        fprintf(stderr, "Error: %s %s %d\n", cudaGetErrorString(code), file, line);
        exit(code);
    }
}
#endif

namespace mpr {

// inc/gpu_opcode.hpp: numeric values are part of the tape format
enum Opcode {
    GPU_OP_INVALID = 0, GPU_OP_JUMP,
    GPU_OP_SQUARE_LHS, GPU_OP_SQRT_LHS, GPU_OP_NEG_LHS, GPU_OP_SIN_LHS, GPU_OP_COS_LHS, GPU_OP_ASIN_LHS,
    GPU_OP_ACOS_LHS, GPU_OP_ATAN_LHS, GPU_OP_EXP_LHS, GPU_OP_ABS_LHS, GPU_OP_LOG_LHS,
    GPU_OP_ADD_LHS_IMM, GPU_OP_ADD_LHS_RHS, GPU_OP_MUL_LHS_IMM, GPU_OP_MUL_LHS_RHS,
    GPU_OP_MIN_LHS_IMM, GPU_OP_MIN_LHS_RHS, GPU_OP_MAX_LHS_IMM, GPU_OP_MAX_LHS_RHS,
    GPU_OP_SUB_LHS_IMM, GPU_OP_SUB_IMM_RHS, GPU_OP_SUB_LHS_RHS,
    GPU_OP_DIV_LHS_IMM, GPU_OP_DIV_IMM_RHS, GPU_OP_DIV_LHS_RHS,
    GPU_OP_COPY_IMM, GPU_OP_COPY_LHS, GPU_OP_COPY_RHS,
};
const char* gpu_op_str(uint8_t op);

// inc/util.hpp.  Buffers handed out by a Context / Tape stay owned by the underlying
// C handle, so their Ptr<> carries a deleter that is switched off; Ptr<>s made by user code
// (default-constructed deleter) free with mprb_free_device, as the reference's do with cudaFree.
struct Deleter {
    bool owns = true;
    template <typename T> void operator()(T* p) const { if (owns && p) mprb_free_device((void*)p); }
};
template <typename T> using Ptr = std::unique_ptr<T, Deleter>;

inline constexpr unsigned pow(unsigned p, unsigned n) { return n ? p * pow(p, n - 1) : 1; }

namespace detail {
inline void check(int code, const char* what) {      // CUDA_CHECK convention: print and exit
    if (code == MPRB_E_OVERFLOW) {
        // A tile list outgrew its array (capped at 64 Mi tiles per level; the reference would have
        // reallocated).  The frame is incomplete but the context is intact: say so and carry on,
        // like the reference does when its subtape arena runs out (context.cu:336-347).
        fprintf(stderr, "Warning: %s: %s\n", what, mprb_last_error());
        return;
    }
    if (code != 0) {
        fprintf(stderr, "Error: %s: %s\n", what, mprb_last_error());
        exit(code);
    }
}
template <typename T> Ptr<T> borrowed(typename Ptr<T>::pointer p) { return Ptr<T>(p, Deleter{false}); }
struct CtxClose { void operator()(mprb_ctx* c) const { mprb_ctx_destroy(c); } };
struct TapeClose { void operator()(mprb_tape* t) const { mprb_tape_destroy(t); } };
}  // namespace detail

// inc/tape.hpp
struct Tape {
    Tape(const libfive::Tree& tree);               // packs the tree (restates src/tape.cpp) and uploads
    Tape(const uint64_t* cells, int32_t n);        // extra: wrap already-packed cells
    Ptr<uint64_t[]> data;                          // GPU (managed) memory
    int32_t length = 0;
    std::unique_ptr<mprb_tape, detail::TapeClose> handle;
};

// inc/context.hpp
struct TileNode { int32_t position; int32_t tape; int32_t next; };

struct Tiles {
    Ptr<int32_t[]> filled;
    Ptr<TileNode[]> tiles;
    size_t tile_array_size = 0;
};

struct Context {
    Context(int32_t image_size_px) : image_size_px(image_size_px) {
        mprb_ctx_opts opts = {-1, NUM_SUBTAPES, 0, 0, 0, 0, 0};
        mprb_ctx* c = nullptr;
        detail::check(mprb_ctx_create(image_size_px, &opts, &c), "mprb_ctx_create");
        handle.reset(c);
        refresh();
    }
    void render3D(const Tape& tape, const Eigen::Matrix4f& mat) {
        detail::check(mprb_render3d(handle.get(), tape.handle.get(), mat.data()), "render3D");
        refresh();
    }
    void render2D(const Tape& tape, const Eigen::Matrix3f& mat, const float z = 0.0f) {
        detail::check(mprb_render2d(handle.get(), tape.handle.get(), mat.data(), z), "render2D");
        refresh();
    }

    void render2D_brute(const Tape& tape, const Eigen::Matrix3f& mat, const float z = 0.0f) {
        detail::check(mprb_render2d_brute(handle.get(), tape.handle.get(), mat.data(), z), "render2D_brute");
        refresh();
    }
    Ptr<float[]> render2D_heatmap(const Tape& tape, const Eigen::Matrix3f& mat, const float z = 0.0f) {
        float* h = nullptr;
        detail::check(mprb_render2d_heatmap(handle.get(), tape.handle.get(), mat.data(), z, &h), "render2D_heatmap");
        refresh();
        return Ptr<float[]>(h);
    }
    Ptr<float[]> render3D_heatmap(const Tape& tape, const Eigen::Matrix4f& mat) {
        float* h = nullptr;
        detail::check(mprb_render3d_heatmap(handle.get(), tape.handle.get(), mat.data(), &h), "render3D_heatmap");
        refresh();
        return Ptr<float[]>(h);
    }

    int32_t image_size_px;
    Ptr<uint64_t[]> tape_data;
    Ptr<int32_t> tape_index;
    Tiles stages[4];
    Ptr<int32_t> num_active_tiles;
    Ptr<void> values;                              // unused here: transforms are fused into the kernels
    size_t values_size = 0;
    Ptr<uint32_t[]> normals;
    std::unique_ptr<mprb_ctx, detail::CtxClose> handle;

private:
    void refresh() {                               // (re)bind the public members to the handle's buffers
        mprb_buffers b;
        detail::check(mprb_ctx_buffers(handle.get(), &b), "mprb_ctx_buffers");
        tape_data = detail::borrowed<uint64_t[]>(b.tape_data);
        tape_index = detail::borrowed<int32_t>(b.tape_index);
        num_active_tiles = detail::borrowed<int32_t>(b.num_active_tiles);
        normals = detail::borrowed<uint32_t[]>(b.normals);
        for (int i = 0; i < 4; ++i) {
            stages[i].filled = detail::borrowed<int32_t[]>(b.filled[i]);
            stages[i].tiles = detail::borrowed<TileNode[]>(reinterpret_cast<TileNode*>(b.tiles[i]));
            stages[i].tile_array_size = size_t(b.tile_array_size[i]);
        }
    }
};

// inc/effects.hpp
struct Effects {
    Effects() {
        // Same draws, same order as the reference constructor (src/effects.cu:229-250): the sample
        // sets depend on the process-wide rand() state exactly as they do there.
        float kernel[64 * 3], rvecs[256 * 3];                       // column-major 64x3 / 256x3
        for (unsigned i = 0; i < 64; ++i) {
            float v[3] = {2.0f * ((float)(rand()) / (float)(RAND_MAX) - 0.5f),
                          2.0f * ((float)(rand()) / (float)(RAND_MAX) - 0.5f),
                          (float)(rand()) / (float)(RAND_MAX)};
            const float n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
            float scale = float(i) / float(64 - 1);
            scale = (scale * scale) * 0.9f + 0.1f;
            for (int k = 0; k < 3; ++k) kernel[k * 64 + i] = v[k] / n * scale;
        }
        for (unsigned i = 0; i < 256; ++i) {
            float v[3] = {2.0f * ((float)(rand()) / (float)(RAND_MAX) - 0.5f),
                          2.0f * ((float)(rand()) / (float)(RAND_MAX) - 0.5f), 0.0f};
            const float n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
            for (int k = 0; k < 3; ++k) rvecs[k * 256 + i] = v[k] / n;
        }
        mprb_effects* fx = nullptr;
        detail::check(mprb_effects_create(kernel, rvecs, &fx), "mprb_effects_create");
        handle.reset(fx, [](mprb_effects* p) { mprb_effects_destroy(p); });
    }
    Ptr<int32_t[]> tmp;
    Ptr<int32_t[]> image;
    void drawSSAO(const Context& ctx) {
        detail::check(mprb_effects_draw_ssao(handle.get(), ctx.handle.get()), "drawSSAO");
        refresh();
    }
    void drawShaded(const Context& ctx) {
        detail::check(mprb_effects_draw_shaded(handle.get(), ctx.handle.get()), "drawShaded");
        refresh();
    }
    std::shared_ptr<mprb_effects> handle;

private:
    void refresh() {
        int32_t *i = nullptr, *t = nullptr;
        detail::check(mprb_effects_buffers(handle.get(), &i, &t), "mprb_effects_buffers");
        image = detail::borrowed<int32_t[]>(i);
        tmp = detail::borrowed<int32_t[]>(t);
    }
};

}  // namespace mpr
