// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// A C-ABI shim around the UNMODIFIED reference renderer so that tests and the
// bench's reference arm can drive it with ctypes.  It is compiled together
// with the reference's own sources *where they lie* under /root/reference
// (src/context.cu, src/context.cpp, src/gpu_opcode.cu + inc/*.hpp) by
// oracle/Makefile into oracle/_ref/libmpr_ref.so; no reference source is
// copied into this repository.
//
// The reference builds mpr::Tape from a libfive::Tree (src/tape.cpp), and
// libfive cannot be built here (Eigen/Boost/libpng are absent).  tape.hpp only
// forward-declares libfive::Tree, so this file supplies a carrier type with
// that name which simply holds an already-packed clause array, plus the
// Tape constructor that uploads it the way src/tape.cpp:223-227 does.
#include <cstdint>
#include <cstring>

#include "context.hpp"
#include "effects.hpp"
#include "parameters.hpp"
#include "tape.hpp"

namespace libfive {
class Tree {
public:
    const uint64_t* cells;
    int32_t n;
};
}  // namespace libfive

namespace mpr {
Tape::Tape(const libfive::Tree& t) {
    data.reset(CUDA_MALLOC(uint64_t, t.n));
    CUDA_CHECK(cudaMemcpy(data.get(), t.cells, sizeof(uint64_t) * t.n,
                          cudaMemcpyHostToDevice));
    length = t.n;
}
}  // namespace mpr

extern "C" {

void* ref_ctx_create(int image_size_px) { return new mpr::Context(image_size_px); }
void ref_ctx_destroy(void* c) { delete static_cast<mpr::Context*>(c); }

void* ref_tape_create(const uint64_t* host_cells, int32_t n) {
    libfive::Tree t;
    t.cells = host_cells;
    t.n = n;
    return new mpr::Tape(t);
}
void ref_tape_destroy(void* t) { delete static_cast<mpr::Tape*>(t); }

void ref_render2d(void* c, void* t, const float* mat3_colmajor, float z) {
    Eigen::Matrix3f m;
    memcpy(m.d, mat3_colmajor, sizeof(float) * 9);
    static_cast<mpr::Context*>(c)->render2D(*static_cast<mpr::Tape*>(t), m, z);
}
void ref_render3d(void* c, void* t, const float* mat4_colmajor) {
    Eigen::Matrix4f m;
    memcpy(m.d, mat4_colmajor, sizeof(float) * 16);
    static_cast<mpr::Context*>(c)->render3D(*static_cast<mpr::Tape*>(t), m);
}

void ref_render2d_brute(void* c, void* t, const float* mat3_colmajor, float z) {
    Eigen::Matrix3f m;
    memcpy(m.d, mat3_colmajor, sizeof(float) * 9);
    static_cast<mpr::Context*>(c)->render2D_brute(*static_cast<mpr::Tape*>(t), m, z);
}
// The heatmap variants return a managed S*S float array owned by the caller; it is copied out
// and released here.
void ref_render2d_heatmap(void* c, void* t, const float* mat3_colmajor, float z, float* heat_out) {
    Eigen::Matrix3f m;
    memcpy(m.d, mat3_colmajor, sizeof(float) * 9);
    mpr::Context* ctx = static_cast<mpr::Context*>(c);
    auto h = ctx->render2D_heatmap(*static_cast<mpr::Tape*>(t), m, z);
    memcpy(heat_out, h.get(), sizeof(float) * size_t(ctx->image_size_px) * ctx->image_size_px);
}
void ref_render3d_heatmap(void* c, void* t, const float* mat4_colmajor, float* heat_out) {
    Eigen::Matrix4f m;
    memcpy(m.d, mat4_colmajor, sizeof(float) * 16);
    mpr::Context* ctx = static_cast<mpr::Context*>(c);
    auto h = ctx->render3D_heatmap(*static_cast<mpr::Tape*>(t), m);
    memcpy(heat_out, h.get(), sizeof(float) * size_t(ctx->image_size_px) * ctx->image_size_px);
}

int32_t* ref_filled(void* c, int stage) {
    return static_cast<mpr::Context*>(c)->stages[stage].filled.get();
}
void* ref_tiles(void* c, int stage) {
    return static_cast<mpr::Context*>(c)->stages[stage].tiles.get();
}
uint64_t ref_tile_array_size(void* c, int stage) {
    return static_cast<mpr::Context*>(c)->stages[stage].tile_array_size;
}
uint64_t* ref_tape_data(void* c) { return static_cast<mpr::Context*>(c)->tape_data.get(); }
int32_t ref_tape_index(void* c) { return *static_cast<mpr::Context*>(c)->tape_index; }
int32_t ref_num_active(void* c) { return *static_cast<mpr::Context*>(c)->num_active_tiles; }
uint32_t* ref_normals(void* c) { return static_cast<mpr::Context*>(c)->normals.get(); }
int64_t ref_num_subtapes(void) { return NUM_SUBTAPES; }

// Bench helper: copies the final image (and normals) to host buffers with cudaMemcpy, the
// cheapest way to read the reference's managed result buffers from the host.
void ref_download(void* c, int32_t* image_out, uint32_t* normals_out) {
    mpr::Context* ctx = static_cast<mpr::Context*>(c);
    const size_t n = size_t(ctx->image_size_px) * ctx->image_size_px;
    if (image_out) cudaMemcpy(image_out, ctx->stages[3].filled.get(), n * sizeof(int32_t), cudaMemcpyDeviceToHost);
    if (normals_out) cudaMemcpy(normals_out, ctx->normals.get(), n * sizeof(uint32_t), cudaMemcpyDeviceToHost);
}

// ---- mpr::Effects (src/effects.cu), compiled unmodified against oracle/shim/Eigen ----------------
// The sample sets are protected members filled from rand() by the constructor; a derived type
// exposes them so that the product can be handed the very same numbers.
struct RefEffects : mpr::Effects {
    const float* kernel() const { return ssao_kernel.data(); }       // 64 x 3, column major
    const float* rvecs() const { return ssao_rvecs.data(); }         // 256 x 3, column major
};
void* ref_effects_create(void) { return new RefEffects(); }
void ref_effects_destroy(void* fx) { delete static_cast<RefEffects*>(fx); }
void ref_effects_samples(void* fx, float* kernel_64x3, float* rvecs_256x3) {
    memcpy(kernel_64x3, static_cast<RefEffects*>(fx)->kernel(), sizeof(float) * 64 * 3);
    memcpy(rvecs_256x3, static_cast<RefEffects*>(fx)->rvecs(), sizeof(float) * 256 * 3);
}
void ref_effects_draw_ssao(void* fx, void* c) { static_cast<RefEffects*>(fx)->drawSSAO(*static_cast<mpr::Context*>(c)); }
void ref_effects_draw_shaded(void* fx, void* c) { static_cast<RefEffects*>(fx)->drawShaded(*static_cast<mpr::Context*>(c)); }
int32_t* ref_effects_image(void* fx) { return static_cast<RefEffects*>(fx)->image.get(); }
int32_t* ref_effects_tmp(void* fx) { return static_cast<RefEffects*>(fx)->tmp.get(); }

}  // extern "C"
