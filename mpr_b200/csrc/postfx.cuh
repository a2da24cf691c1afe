// Launchers of the post-effect kernels (postfx.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mprb {
void launch_clear_pair(int32_t* a, int32_t* b, long long n, cudaStream_t s);   // n: a multiple of 4
void launch_draw_ssao(const int32_t* depth, const uint32_t* norm, const float* kernel, const float* rvecs,
                      int size, int32_t* out, cudaStream_t s);
void launch_blur_ssao(const int32_t* image, const int32_t* ssao, int size, int32_t* out, cudaStream_t s);
void launch_draw_shaded(const int32_t* depth, const uint32_t* norm, const int32_t* ssao, int size, int32_t* out,
                        cudaStream_t s);
}  // namespace mprb
