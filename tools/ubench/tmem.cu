// Micro-benchmark (not part of the library): tensor memory as plain per-lane scratch.  One warp stores
// and loads float2 "rows" at dynamic columns with tcgen05.st / tcgen05.ld (32x32b.x2: lane i of the warp
// touches TMEM lane 32 * (warp % 4) + i, two 32-bit columns) and checks the values; reports cycles per
// dependent store -> load round trip and per independent load.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/ubench/tmem.cu -o build/ubench_tmem
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void tm_st2(uint32_t taddr, uint32_t a, uint32_t b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void tm_ld2(uint32_t taddr, uint32_t& a, uint32_t& b) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void k(long long* out, int* bad, int reps) {
    __shared__ uint32_t base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"l"((unsigned long long)__cvta_generic_to_shared(&base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = base_s + ((uint32_t(warp & 3) * 32u) << 16);      // lane field in bits 16+
    // correctness: rows r = 0..15 hold (1000 * warp + 32 * r + lane, ~that)
    for (int r = 0; r < 16; ++r) tm_st2(base + 2 * r, 1000 * warp + 32 * r + lane, ~(1000 * warp + 32 * r + lane));
    tm_wait_st();
    int errors = 0;
    for (int r = 15; r >= 0; --r) {
        uint32_t a, b;
        tm_ld2(base + 2 * r, a, b);
        tm_wait_ld();
        if (a != uint32_t(1000 * warp + 32 * r + lane) || b != ~a) ++errors;
    }
    if (errors) atomicAdd(bad, errors);
    // dependent chain: value -> st -> wait -> ld (another row index derived from the value) -> wait
    uint32_t v = lane, w = 0, col = 0;
    long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
        tm_st2(base + col, v, w);
        tm_wait_st();
        tm_ld2(base + col, v, w);
        tm_wait_ld();
        col = (v + i) & 30;          // data-dependent next address
        v += 1;
    }
    long long t1 = clock64();
    // independent loads (throughput of ld + wait pairs with no store in between)
    uint32_t acc = 0;
    long long t2 = clock64();
    for (int i = 0; i < reps; ++i) {
        uint32_t a, b;
        tm_ld2(base + ((i * 2) & 30), a, b);
        tm_wait_ld();
        acc += a;
    }
    long long t3 = clock64();
    if (lane == 0) { out[warp * 2] = t1 - t0; out[warp * 2 + 1] = t3 - t2; }
    if (acc == 0xdeadbeef) out[100] = v;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(base_s) : "memory");
}

int main() {
    long long* out; int* bad;
    cudaMalloc(&out, 1024); cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
    const int reps = 10000;
    k<<<1, 128>>>(out, bad, reps);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[8]; int hb = 0;
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
    printf("status: %s, value errors: %d\n", cudaGetErrorString(e), hb);
    for (int w = 0; w < 4; ++w)
        printf("warp %d: st+wait+ld+wait round trip %.1f cycles, ld+wait %.1f cycles\n", w, double(h[2 * w]) / reps, double(h[2 * w + 1]) / reps);
    return 0;
}
