// Micro-benchmark of the float pass's clause loop (not part of the library): cycles per clause of the
// generated PTX loops for one warp alone (the dependent chain) and for W warps per SM (throughput).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I mpr_b200/csrc tools/ubench/dispatch.cu -o build/ubench_dispatch
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <int G, int U>
__device__ __forceinline__ void run(uint32_t& cp, uint32_t& w, uint32_t& imm, uint32_t sb);
#define LOOP(G, U, FILE)                                                                               \
    template <> __device__ __forceinline__ void run<G, U>(uint32_t & cp, uint32_t & w, uint32_t & imm, uint32_t sb) { \
        asm volatile(
#define LOOP_END : "+r"(cp), "=&r"(w), "=&r"(imm) : "r"(sb) : "memory"); }
template <> __device__ __forceinline__ void run<1, 1>(uint32_t& cp, uint32_t& w, uint32_t& imm, uint32_t sb) {
    asm volatile(
#include "float_loop_ptx.inc"
        : "+r"(cp), "=&r"(w), "=&r"(imm) : "r"(sb) : "memory");
}
template <> __device__ __forceinline__ void run<2, 1>(uint32_t& cp, uint32_t& w, uint32_t& imm, uint32_t sb) {
    asm volatile(
#include "float_loop_ptx_g2.inc"
        : "+r"(cp), "=&r"(w), "=&r"(imm) : "r"(sb) : "memory");
}
template <> __device__ __forceinline__ void run<4, 1>(uint32_t& cp, uint32_t& w, uint32_t& imm, uint32_t sb) {
    asm volatile(
#include "float_loop_ptx_g4.inc"
        : "+r"(cp), "=&r"(w), "=&r"(imm) : "r"(sb) : "memory");
}
// smem per warp: 64 cells (512 B) + 32 slot rows of 256*G bytes
template <int G, int U>
__global__ void k(const uint64_t* tape, int n_cells, int reps, long long* cycles, float* sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int per_warp = 1024 + 32 * 256 * G;
    unsigned char* mine = sm + warp * per_warp;
    uint64_t* cells = reinterpret_cast<uint64_t*>(mine);
    for (int i = lane; i < 128; i += 32) cells[i] = i < n_cells ? tape[i] : 0;   // cell n_cells.. = END
    float* rows = reinterpret_cast<float*>(mine + 1024);
    for (int i = lane; i < 32 * 64 * G; i += 32) rows[i] = 1.0f + 0.001f * i;
    __syncwarp();
    const uint32_t base = uint32_t(__cvta_generic_to_shared(mine));
    const uint32_t sb = base + 1024 + lane * 8 * G;
    long long t0 = clock64();
    uint32_t w = 0, imm = 0;
    for (int r = 0; r < reps; ++r) {
        uint32_t cp = base;           // cell 0 is a dummy header; the loop starts at cell 1
        run<G, U>(cp, w, imm, sb);
    }
    long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
    if (w == 12345) sink[0] = rows[lane];
}

static uint64_t cell(int op, int out, int lhs, int rhs, float imm, int G, int hints) {
    uint32_t ib; memcpy(&ib, &imm, 4);
    return uint64_t(uint32_t(op | hints) | (uint32_t(out * G) << 8) | (uint32_t(lhs * G) << 16) | (uint32_t(rhs * G) << 24)) | (uint64_t(ib) << 32);
}

template <int G, int U> void bench(const char* name, int kind, int warps, int ctas_per_sm) {
    const int n = 62;
    std::vector<uint64_t> t(64, 0);
    for (int i = 1; i <= n; ++i) {
        if (kind == 0) t[i] = cell(13, 5, 5, 0, 0.5f, G, (i > 1 ? 0x20 : 0) | (i < n ? 0x80 : 0));  // ADD_LI, forwarded, no store
        else if (kind == 1) t[i] = cell(14, 1 + (i % 7), 8 + (i % 5), 16 + (i % 3), 0.f, G, 0);       // ADD_LR, loads + store
        else t[i] = cell((i & 1) ? 16 : 13, 1 + (i % 7), (i > 1) ? 1 + ((i - 1) % 7) : 9, 16 + (i % 3), 1.5f, G, (i > 1 ? 0x20 : 0)); // fwd lhs, store
    }
    uint64_t* d; cudaMalloc(&d, 64 * 8); cudaMemcpy(d, t.data(), 64 * 8, cudaMemcpyHostToDevice);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * ctas_per_sm, reps = 2000;
    long long* cyc; cudaMalloc(&cyc, sizeof(long long) * grid * warps);
    float* sink; cudaMalloc(&sink, 4);
    const size_t smem = size_t(warps) * (1024 + 32 * 256 * G);
    cudaFuncSetAttribute(k<G, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    k<G, U><<<grid, warps * 32, smem>>>(d, n + 1, 10, cyc, sink);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<G, U><<<grid, warps * 32, smem>>>(d, n + 1, reps, cyc, sink);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(grid * warps); cudaMemcpy(h.data(), cyc, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
    double avg = 0; for (auto v : h) avg += double(v); avg /= h.size();
    const cudaError_t err = cudaGetLastError();
    printf("%-28s G=%d U=%d warps/SM=%3d  cycles/clause/warp=%7.1f  tile-clauses/cycle/SM=%6.3f  (%s)\n", name, G, U,
           warps * ctas_per_sm, avg / (double(reps) * n), double(warps * ctas_per_sm) * G / (avg / (double(reps) * n)),
           cudaGetErrorString(err));
    cudaFree(d); cudaFree(cyc); cudaFree(sink);
}

int main() {
    const char* names[3] = {"ADD_LI fwd+nostore (head)", "ADD_LR 2 loads + store", "mixed fwd-lhs + store"};
    for (int kind = 0; kind < 3; ++kind) {
        bench<1, 1>(names[kind], kind, 1, 1);
        bench<1, 1>(names[kind], kind, 17, 2);
        bench<2, 1>(names[kind], kind, 1, 1);
        bench<2, 1>(names[kind], kind, 12, 1);
        bench<4, 1>(names[kind], kind, 1, 1);
        bench<4, 1>(names[kind], kind, 6, 1);
    }
    return 0;
}
