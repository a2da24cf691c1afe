// Post-effects on a rendered 3D frame: screen-space ambient occlusion, a variance-guided
// blur, and single-light shading (reference src/effects.cu:17-297, inc/effects.hpp:21-37).
//
// The reference passes two Eigen matrices by value as kernel arguments (64x3 hemisphere
// samples, 16x16 x 3 rotation vectors = 3840 bytes); here they live in global memory and are
// read through the read-only path.  One thread per pixel, 16x16 blocks, as in the reference:
// the per-pixel rotation vector is indexed by (threadIdx.x % 16) * 16 + threadIdx.y % 16.
//
// PARITY: src/effects.cu compiles unmodified into the test oracle (oracle/_ref, with a stand-in for
// Eigen's small fixed-size vectors in oracle/shim/Eigen/Eigen), and tests/test_gpu_parity.py compares
// both result buffers of drawSSAO / drawShaded with it bit for bit.  The vector helpers below use
// the operation order of Eigen 3.3 for these sizes - linear reductions (a0*b0 + a1*b1) + a2*b2,
// normalized() = v / sqrt(v.v) when v.v > 0, coefficient-wise 3x3 * vector - and this file is built
// with the compiler's default floating-point contraction, like the reference.  The source's quirks are
// kept: the second loop of blur_ssao samples around the image origin (effects.cu:131-132), and the
// bounds tests of draw_ssao / draw_shaded use `&&` (effects.cu:31, :175), which lets nothing out as
// long as the image side is a multiple of the 16-px block (it always is here: sizes are multiples of 64).
#include <cstdint>
#include <cuda_runtime.h>

#include "postfx.cuh"

namespace mprb {

namespace {

struct V3 { float d[3]; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.d[0] = x; r.d[1] = y; r.d[2] = z; return r; }
__device__ __forceinline__ float dot(const V3& a, const V3& b) {
    float s = a.d[0] * b.d[0];
    #pragma unroll
    for (int i = 1; i < 3; ++i) s = s + a.d[i] * b.d[i];
    return s;
}
__device__ __forceinline__ V3 scale(const V3& a, float s) {
    V3 r;
    #pragma unroll
    for (int i = 0; i < 3; ++i) r.d[i] = a.d[i] * s;
    return r;
}
__device__ __forceinline__ V3 quot(const V3& a, float s) {
    V3 r;
    #pragma unroll
    for (int i = 0; i < 3; ++i) r.d[i] = a.d[i] / s;
    return r;
}
__device__ __forceinline__ V3 sub(const V3& a, const V3& b) {
    V3 r;
    #pragma unroll
    for (int i = 0; i < 3; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
__device__ __forceinline__ V3 add(const V3& a, const V3& b) {
    V3 r;
    #pragma unroll
    for (int i = 0; i < 3; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
__device__ __forceinline__ V3 cross(const V3& a, const V3& b) {
    return v3(a.d[1] * b.d[2] - a.d[2] * b.d[1], a.d[2] * b.d[0] - a.d[0] * b.d[2], a.d[0] * b.d[1] - a.d[1] * b.d[0]);
}
// Eigen's normalized(): divide by the norm unless the squared norm is zero
__device__ __forceinline__ V3 normalized(const V3& a) {
    const float z = dot(a, a);
    if (z > 0.0f) return quot(a, sqrtf(z));
    return a;
}

__global__ void k_draw_ssao(const int32_t* __restrict__ depth, const uint32_t* __restrict__ norm,
                            const float* __restrict__ kernel /* 64x3 col-major */,
                            const float* __restrict__ rvecs /* 256x3 col-major */,
                            int size, int32_t* __restrict__ output)
{
    const int x = threadIdx.x + blockIdx.x * blockDim.x;
    const int y = threadIdx.y + blockIdx.y * blockDim.y;
    constexpr float RADIUS = 0.1f;
    if (x >= size && y >= size) return;                    // sic (effects.cu:31)
    const int h = depth[x + y * size];
    if (!h) return;
    const float3 pos = make_float3(2.0f * ((x + 0.5f) / size - 0.5f), 2.0f * ((y + 0.5f) / size - 0.5f),
                                   2.0f * ((h + 0.5f) / size - 0.5f));
    const uint32_t n = norm[x + y * size];
    const float dx = (float)(n & 0xFF) - 128.0f;
    const float dy = (float)((n >> 8) & 0xFF) - 128.0f;
    const float dz0 = (float)((n >> 16) & 0xFF) - 128.0f;
    const V3 normal = normalized(v3(dx, dy, dz0));
    const int ri = (threadIdx.x % 16) * 16 + (threadIdx.y % 16);
    const V3 rvec = v3(__ldg(&rvecs[ri]), __ldg(&rvecs[256 + ri]), __ldg(&rvecs[512 + ri]));
    const V3 tangent = normalized(sub(rvec, scale(normal, dot(rvec, normal))));
    const V3 bitangent = cross(normal, tangent);
    const V3 p3 = v3(pos.x, pos.y, pos.z);

    float occlusion = 0.0f;
    for (unsigned i = 0; i < 64; ++i) {
        const V3 k = v3(__ldg(&kernel[i]), __ldg(&kernel[64 + i]), __ldg(&kernel[128 + i]));
        // tbn * k with columns (tangent, bitangent, normal): row r = (t[r] * k0 + b[r] * k1) + n[r] * k2
        V3 r;
        #pragma unroll
        for (int c = 0; c < 3; ++c) {
            float sum = tangent.d[c] * k.d[0];
            sum = sum + bitangent.d[c] * k.d[1];
            sum = sum + normal.d[c] * k.d[2];
            r.d[c] = sum;
        }
        const V3 sp = add(scale(r, RADIUS), p3);
        const unsigned px = (sp.d[0] / 2.0f + 0.5f) * size;
        const unsigned py = (sp.d[1] / 2.0f + 0.5f) * size;
        const unsigned actual_h = (px < size && py < size) ? depth[px + py * size] : 0;
        const float actual_z = 2.0f * ((actual_h + 0.5f) / size - 0.5f);
        const auto dz = fabsf(sp.d[2] - actual_z);
        if (dz < RADIUS) {
            occlusion += sp.d[2] <= actual_z;
        } else if (dz < RADIUS * 2.0f) {
            if (sp.d[2] <= actual_z) occlusion += powf((RADIUS - (dz - RADIUS)) / RADIUS, 2.0f);
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
    if (x >= size && y >= size) return;                    // sic (effects.cu:175)
    const auto h = depth[x + y * size];
    if (!h) return;
    const uint8_t s = ssao[x + y * size];
    const auto n = norm[x + y * size];
    float dx = (float)(n & 0xFF) - 128.0f;
    float dy = (float)((n >> 8) & 0xFF) - 128.0f;
    float dz = (float)((n >> 16) & 0xFF) - 128.0f;
    const V3 normal = normalized(v3(dx, dy, dz));
    const float3 pos_f3 = make_float3(2.0f * ((x + 0.5f) / size - 0.5f), 2.0f * ((y + 0.5f) / size - 0.5f),
                                      2.0f * ((h + 0.5f) / size - 0.5f));
    const V3 pos = v3(pos_f3.x, pos_f3.y, pos_f3.z);
    const V3 light_pos = v3(5, 5, 10);
    const V3 light_dir = normalized(sub(light_pos, pos));
    float light = fmaxf(0.0f, dot(light_dir, normal)) * 0.8f;
    light *= s / 255.0f;
    light += 0.2f;
    if (light < 0.0f) light = 0.0f;
    else if (light > 1.0f) light = 1.0f;
    uint8_t color = light * 255.0f;
    output[x + y * size] = (0xFF << 24) | (color << 16) | (color << 8) | (color << 0);
}

// Both result images to zero in one launch (the reference memsets them, effects.cu:265-266 / :283-284; they
// are managed memory here and a driver memset on a managed range stalls the host, see k_begin_frame).
__global__ void k_clear_pair(uint4* __restrict__ a, uint4* __restrict__ b, long long n16)
{
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
        a[i] = zero;
        b[i] = zero;
    }
}

}  // namespace

void launch_clear_pair(int32_t* a, int32_t* b, long long n, cudaStream_t s) {
    const long long n16 = n / 4;                    // image sizes are multiples of 64 px
    const int grid = int(n16 / 256 < 1 ? 1 : (n16 / 256 > 148 * 16 ? 148 * 16 : n16 / 256));
    k_clear_pair<<<grid, 256, 0, s>>>(reinterpret_cast<uint4*>(a), reinterpret_cast<uint4*>(b), n16);
}
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
