// Post-effects on a rendered 3D frame: screen-space ambient occlusion, a variance-guided
// blur, and single-light shading (reference src/effects.cu:17-297, inc/effects.hpp:21-37).
//
// The reference passes two Eigen matrices by value as kernel arguments (64x3 hemisphere
// samples, 16x16 x 3 rotation vectors = 3840 bytes); here they live in global memory and are
// read through the read-only path.  One thread per pixel, 16x16 blocks, as in the reference:
// the per-pixel rotation vector is indexed by (threadIdx.x % 16) * 16 + threadIdx.y % 16.
//
// PARITY: the reference's effects.cu needs real Eigen in device code and cannot be built in
// this environment, so there is no reference build to compare with; this file and the CPU
// restatement in oracle/mpr_oracle.c follow the source text, including its quirks (the second
// loop of blur_ssao samples around the image origin, effects.cu:131-132; the `&&` in the bounds
// test of draw_ssao / draw_shaded).
#include <cstdint>
#include <cuda_runtime.h>

#include "postfx.cuh"

namespace mprb {

namespace {

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 scale(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// Eigen's normalized(): divide by the norm unless the squared norm is zero
__device__ __forceinline__ V3 normalized(V3 a) {
    const float z = dot(a, a);
    if (z > 0.0f) {
        const float n = sqrtf(z);
        return v3(a.x / n, a.y / n, a.z / n);
    }
    return a;
}
__device__ __forceinline__ float ndc(float p, int size) { return 2.0f * ((p + 0.5f) / size - 0.5f); }

__global__ void k_draw_ssao(const int32_t* __restrict__ depth, const uint32_t* __restrict__ norm,
                            const float* __restrict__ kernel /* 64x3 col-major */,
                            const float* __restrict__ rvecs /* 256x3 col-major */,
                            int size, int32_t* __restrict__ output)
{
    const int x = threadIdx.x + blockIdx.x * blockDim.x;
    const int y = threadIdx.y + blockIdx.y * blockDim.y;
    constexpr float RADIUS = 0.1f;
    if (x >= size || y >= size) return;
    const int h = depth[x + y * size];
    if (!h) return;
    const V3 pos = v3(ndc(float(x), size), ndc(float(y), size), ndc(float(h), size));
    const uint32_t n = norm[x + y * size];
    const V3 normal = normalized(v3(float(n & 0xFF) - 128.0f, float((n >> 8) & 0xFF) - 128.0f,
                                    float((n >> 16) & 0xFF) - 128.0f));
    const int ri = (threadIdx.x % 16) * 16 + (threadIdx.y % 16);
    const V3 rvec = v3(__ldg(&rvecs[ri]), __ldg(&rvecs[256 + ri]), __ldg(&rvecs[512 + ri]));
    const V3 tangent = normalized(sub(rvec, scale(normal, dot(rvec, normal))));
    const V3 bitangent = cross(normal, tangent);

    float occlusion = 0.0f;
    for (int i = 0; i < 64; ++i) {
        const V3 k = v3(__ldg(&kernel[i]), __ldg(&kernel[64 + i]), __ldg(&kernel[128 + i]));
        // tbn * k, columns (tangent, bitangent, normal)
        const V3 r = v3(tangent.x * k.x + bitangent.x * k.y + normal.x * k.z,
                        tangent.y * k.x + bitangent.y * k.y + normal.y * k.z,
                        tangent.z * k.x + bitangent.z * k.y + normal.z * k.z);
        const V3 sp = add(scale(r, RADIUS), pos);
        const unsigned px = (sp.x / 2.0f + 0.5f) * size;
        const unsigned py = (sp.y / 2.0f + 0.5f) * size;
        const unsigned actual_h = (px < unsigned(size) && py < unsigned(size)) ? depth[px + py * size] : 0;
        const float actual_z = 2.0f * ((actual_h + 0.5f) / size - 0.5f);
        const float dz = fabsf(sp.z - actual_z);
        if (dz < RADIUS) {
            occlusion += sp.z <= actual_z;
        } else if (dz < RADIUS * 2.0f) {
            if (sp.z <= actual_z) occlusion += powf((RADIUS - (dz - RADIUS)) / RADIUS, 2.0f);
        }
    }
    occlusion = 1.0 - (occlusion / 64);
    const uint8_t o = occlusion * 255;
    output[x + y * size] = o;
}

__global__ void k_blur_ssao(const int32_t* __restrict__ image, const int32_t* __restrict__ ssao,
                            int size, int32_t* __restrict__ output)
{
    const unsigned x = threadIdx.x + blockIdx.x * blockDim.x;
    const unsigned y = threadIdx.y + blockIdx.y * blockDim.y;
    if (x >= unsigned(size) || y >= unsigned(size)) return;
    constexpr int R = 2;
    float best = 1000000.0f;
    float value = 0.0f;
    for (unsigned q = 0; q < 4; ++q) {
        const int xmin = (q & 1) ? 0 : -R;
        const int ymin = (q & 2) ? 0 : -R;
        float sum = 0.0f, count = 0.0f;
        for (int i = 0; i <= R; ++i)
            for (int j = 0; j <= R; ++j) {
                const int tx = int(x) + xmin + i, ty = int(y) + ymin + j;
                if (tx >= 0 && tx < size && ty >= 0 && ty < size && image[tx + ty * size]) {
                    sum += ssao[tx + ty * size];
                    count++;
                }
            }
        const float mean = sum / count;
        float stdev = 0.0f;
        for (int i = 0; i <= R; ++i)
            for (int j = 0; j <= R; ++j) {
                const int tx = xmin + i, ty = ymin + j;      // sic: relative to the origin (effects.cu:131-132)
                if (tx >= 0 && tx < size && ty >= 0 && ty < size && image[tx + ty * size]) {
                    const float d = mean - ssao[tx + ty * size];
                    stdev += d * d;
                }
            }
        stdev /= count - 1.0f;
        stdev = sqrtf(stdev);
        if (stdev < best) {
            best = stdev;
            value = mean;
        }
    }
    output[x + y * size] = value;
}

__global__ void k_draw_shaded(const int32_t* __restrict__ depth, const uint32_t* __restrict__ norm,
                              const int32_t* __restrict__ ssao, int size, int32_t* __restrict__ output)
{
    const unsigned x = threadIdx.x + blockIdx.x * blockDim.x;
    const unsigned y = threadIdx.y + blockIdx.y * blockDim.y;
    if (x >= unsigned(size) || y >= unsigned(size)) return;
    const int h = depth[x + y * size];
    if (!h) return;
    const uint8_t s = ssao[x + y * size];
    const uint32_t n = norm[x + y * size];
    const V3 normal = normalized(v3(float(n & 0xFF) - 128.0f, float((n >> 8) & 0xFF) - 128.0f,
                                    float((n >> 16) & 0xFF) - 128.0f));
    const V3 pos = v3(ndc(float(x), size), ndc(float(y), size), ndc(float(h), size));
    const V3 light_dir = normalized(sub(v3(5.0f, 5.0f, 10.0f), pos));
    float light = fmaxf(0.0f, dot(light_dir, normal)) * 0.8f;
    light *= s / 255.0f;
    light += 0.2f;
    if (light < 0.0f) light = 0.0f;
    else if (light > 1.0f) light = 1.0f;
    const uint8_t color = light * 255.0f;
    output[x + y * size] = (0xFF << 24) | (color << 16) | (color << 8) | (color << 0);
}

}  // namespace

void launch_draw_ssao(const int32_t* depth, const uint32_t* norm, const float* kernel, const float* rvecs,
                      int size, int32_t* out, cudaStream_t s) {
    const unsigned u = (size + 15) / 16;
    k_draw_ssao<<<dim3(u, u), dim3(16, 16), 0, s>>>(depth, norm, kernel, rvecs, size, out);
}
void launch_blur_ssao(const int32_t* image, const int32_t* ssao, int size, int32_t* out, cudaStream_t s) {
    const unsigned u = (size + 15) / 16;
    k_blur_ssao<<<dim3(u, u), dim3(16, 16), 0, s>>>(image, ssao, size, out);
}
void launch_draw_shaded(const int32_t* depth, const uint32_t* norm, const int32_t* ssao, int size, int32_t* out,
                        cudaStream_t s) {
    const unsigned u = (size + 15) / 16;
    k_draw_shaded<<<dim3(u, u), dim3(16, 16), 0, s>>>(depth, norm, ssao, size, out);
}

}  // namespace mprb
